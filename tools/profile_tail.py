"""One rollout + two IPPOLearner.train() calls at the bench shape and nothing else, for
``ncu --set full -k regex:tail_fused -s 2 -c 1 python tools/profile_tail.py`` (a capture costs seconds).
Prints the update phases' CUDA-event times (not valid under ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iplan_b200.runners.synthetic_runner import build_system   # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    sysm = build_system(n_envs=B, env="highway", hazard=0.01, seed=1)
    batch, *_ = sysm.runner.run()
    for it in range(2):
        sysm.learner.insert_episode_batch(batch)
        sysm.learner.events = []
        sysm.learner.train(sysm.runner.t_env)
        torch.cuda.synchronize()
        ev = sysm.learner.events
        upd = {}
        for (tag, e0), (_, e1) in zip(ev[:-1], ev[1:]):
            if tag != "end":
                upd[tag] = upd.get(tag, 0.0) + e0.elapsed_time(e1)
        print(f"train() #{it}: " + "  ".join(f"{k} {v:.2f} ms" for k, v in upd.items()), flush=True)


if __name__ == "__main__":
    main()
