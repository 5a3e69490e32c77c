"""Host-side profile of the reference-facing numpy API path (bench.py's `e2e` leg): where the time of
ParallelRunner.run_reference_api goes, per API call and per Python function.

    python tools/profile_e2e.py [--envs 512]        (on a GPU box; prints cProfile tables)
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iplan_b200.runners.synthetic_runner import build_system            # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=512)
    a = ap.parse_args()
    sysm = build_system(n_envs=a.envs, env="highway", hazard=0.01, seed=112358)
    sysm.run_and_train(api=True)
    sysm.run_and_train(api=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sysm.runner.run()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sysm.runner.run_reference_api()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"device-resident run(): {(t1 - t0) * 1e3:.1f} ms; run_reference_api(): {(t2 - t1) * 1e3:.1f} ms")
    pr = cProfile.Profile()
    pr.enable()
    batch, *_ = sysm.runner.run_reference_api()
    torch.cuda.synchronize()
    pr.disable()
    t3 = time.perf_counter()
    sysm.learner.insert_episode_batch(batch)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    sysm.learner.train(sysm.runner.t_env)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    print(f"insert_episode_batch: {(t4 - t3) * 1e3:.1f} ms; train: {(t5 - t4) * 1e3:.1f} ms")
    for key in ("cumulative", "tottime"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).strip_dirs().sort_stats(key).print_stats(40)
        print(s.getvalue())
