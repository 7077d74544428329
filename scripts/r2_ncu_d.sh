#!/bin/bash
# Round 2, profiler call for the fused diagonal LNCDE step at B = 32 and B = 2048 (ncu only; after a plain run that exited 0).
set -u
O=gpurun_out
mkdir -p $O
CMD="python scripts/variant_check.py --no-oracle --no-parity --only k2"
$CMD > $O/r2f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file $O/r2b_launches_dlncde.csv $CMD > $O/r2f_l.log 2>&1
tail -c 600 $O/r2f_plain.log
