#!/bin/bash
# Reproduces the round's profiles/ evidence on a B200 (run from the repo root, e.g. under gpurun).
#   bash tools/capture_profiles.sh [outdir]
# 1. launch list of one bench step (cold-cache, serialised times: compare SHARES with bench.py's CUDA-event numbers)
# 2. ncu --set full of the two K1 kernels at the bench shape
# 3. ncu --set full of the tcgen05 fc1 forward / backward at the bench shape (the 13th launch of tools/check_fc1_tc5.py)
# then: ncu -i <rep> --page raw --csv > profiles/<name>.csv   (done on the build host; see profiles/README.md)
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:gat_ -s 6 -c 2 -f -o "$OUT/k1_full" \
    python tools/profile_k1.py 512 4 > "$OUT/k1_ncu.log" 2>&1
ncu --set full --clock-control none -k regex:fc1_fwd_tc5 -s 12 -c 1 -f -o "$OUT/tc5_fwd_full" \
    python tools/check_fc1_tc5.py > "$OUT/tc5_fwd_ncu.log" 2>&1
ncu --set full --clock-control none -k regex:fc1_bwd_tc5 -s 12 -c 1 -f -o "$OUT/tc5_bwd_full" \
    python tools/check_fc1_tc5.py > "$OUT/tc5_bwd_ncu.log" 2>&1

# Round 2, third session:
# 4. K1b (tcgen05) / K1c / the decoder kernel of Behavior_policy.learn, one launch each
ncu --set full --clock-control none --import-source on -k regex:behavior_tc5 -s 60 -c 1 -f -o "$OUT/k1b_tc5_full" \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > "$OUT/k1b_ncu.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:controller_step -s 60 -c 1 -f -o "$OUT/k1c_full" \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > "$OUT/k1c_ncu.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_kernel -s 1 -c 1 -f -o "$OUT/beh_dec_full" \
    python tools/check_beh_learn_tile.py --envs 2 --steps 14 --time-envs 128 > "$OUT/beh_ncu.log" 2>&1
# 5. K1b phase stamps, host-side profile of the numpy API path, the auxiliary learners inside the iteration
IPLAN_BEH_DBG=1 python tools/check_behavior_tc5.py --time-envs 512 > "$OUT/k1b_trace.log" 2>&1
python tools/profile_e2e.py > "$OUT/e2e_profile.log" 2>&1
python bench.py --with-aux --no-e2e --no-cpu-baseline > "$OUT/bench_aux.json" 2> "$OUT/bench_aux.err"
