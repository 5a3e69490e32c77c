#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE — stage the reference's own hot-path modules under oracle/_ref/.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_ref.py        # build container only (needs /root/reference)

The reference (wuxiyang1996/iPLAN) is pure Python, so "building" it means making the files of its hot path importable
where the benchmark runs: the modules that `controllers.dcntrl_controller`, `learners.ippo_learner`,
`components.episode_buffer`, `nova.prediction_policy`, `nova.stable_behavior_policy`, `runners.ippo_parallel_runner` and
`observation_wrapper` pull in (found by importing them from /root/reference and listing what got loaded), plus the YAML
files main.py merges (config/default.yaml, config/envs/*.yaml, config/algs/ippo.yaml).  They are copied UNMODIFIED,
where they lie relative to the reference root, into oracle/_ref/ — git-ignored (never part of the history), NOT
gpurun-ignored (travels to the GPU box next to the built .so).  oracle/_ref/MANIFEST.json lists every file with its
sha256 and the reference commit it came from.

Consumers (only these): `bench.py --impl reference` and its `cpu_baseline` leg (oracle/ref_driver.py, kind "reference"),
and tests/test_gpu_dropin.py (the reference's own ParallelRunner.run body driving the iplan_b200 objects).
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

ROOT_MODULES = ["controllers.dcntrl_controller", "learners.ippo_learner", "components.episode_buffer",
                "components.transforms", "nova.GAT_Net", "nova.behavior_net", "nova.prediction_policy",
                "nova.stable_behavior_policy", "runners.ippo_parallel_runner", "observation_wrapper"]
CONFIGS = ["config/default.yaml", "config/algs/ippo.yaml", "config/envs/highway.yaml", "config/envs/simple_spread_Hetero.yaml"]

TRACE = r"""
import sys, warnings, importlib
warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
sys.path.insert(0, %r)
for m in %r:
    importlib.import_module(m)
for f in sorted(m.__file__ for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and m.__file__.startswith(%r + "/")):
    print(f)
"""


def main():
    if not os.path.isdir(REF):
        print(f"make_ref: {REF} not present (GPU box): keeping the staged oracle/_ref as shipped")
        return 0
    res = subprocess.run([sys.executable, "-c", TRACE % (REF, ROOT_MODULES, REF)], capture_output=True, text=True, cwd="/tmp",
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    if res.returncode != 0:
        print(res.stderr[-2000:])
        raise SystemExit("make_ref: importing the reference's hot-path modules failed")
    files = [os.path.relpath(f, REF) for f in res.stdout.split()] + CONFIGS
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    manifest = {}
    for rel in files:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    commit = None
    for cand in (os.path.join(REF, ".git", "HEAD"),):
        try:
            head = open(cand).read().strip()
            commit = head if not head.startswith("ref:") else open(os.path.join(REF, ".git", head.split()[1])).read().strip()
        except Exception:
            pass
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"reference": "wuxiyang1996/iPLAN", "commit": commit or "961d632 (SURVEY.md)", "files": manifest}, f, indent=1)
    print(f"make_ref: staged {len(files)} reference files under {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
