#!/bin/bash
# Round 2, profiler call for the depth-3 log-signature kernel (ncu only; after a plain run that exited 0).
set -u
O=gpurun_out
mkdir -p $O
python scripts/k1_sweep.py > $O/r2e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logsig_d3 -c 1 -o $O/r2b_full_k1_d3 python scripts/k1_sweep.py > $O/r2e_c.log 2>&1
ncu -i $O/r2b_full_k1_d3.ncu-rep --page raw --csv > $O/r2b_full_k1_d3.raw.csv 2> /dev/null
ncu -i $O/r2b_full_k1_d3.ncu-rep --page source --csv > $O/r2b_src_k1_d3.csv 2> /dev/null
sz=$(stat -c %s $O/r2b_full_k1_d3.ncu-rep 2> /dev/null || echo 0)
if [ "$sz" -gt 6000000 ]; then rm -f $O/r2b_full_k1_d3.ncu-rep; fi
cat $O/r2e_plain.log | tail -7
