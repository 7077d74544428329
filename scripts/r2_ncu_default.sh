#!/bin/bash
# Round 2, profiler call: launch list of the TIMED REGION of the default (EigenWorms Log-NCDE) bench on the final tree -
# bench.py brackets its timed steps with cudaProfilerStart / Stop under LNCDE_BENCH_PROFILE_RANGE=1 (ncu only; after a
# plain run of the same command that exited 0).
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
$B > $O/r2g_plain.log 2>&1 &&
LNCDE_BENCH_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/r2b_launches_logncde.csv $B > $O/r2g_l.log 2>&1
ls -la $O/r2b_launches_logncde.csv
