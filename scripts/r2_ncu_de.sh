#!/bin/bash
# Round 2, profiler call for the dense LNCDE step only (ncu is the only tool of the call; every capture follows a plain
# run of the same command that exited 0).
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload eigenworms_de"
$B > $O/r2d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches_eigenworms_de.csv $B > $O/r2d_l.log 2>&1
$B > $O/r2d_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'scan_fwd_cluster|scan_lam_cluster|dlie_t' -s 6 -c 3 \
    -o $O/r2b_full_k2_de $B > $O/r2d_c.log 2>&1
ncu -i $O/r2b_full_k2_de.ncu-rep --page raw --csv > $O/r2b_full_k2_de.raw.csv 2> /dev/null
ncu -i $O/r2b_full_k2_de.ncu-rep --page source --csv -k regex:scan_lam_cluster > $O/r2b_src_lam_cluster.csv 2> /dev/null
sz=$(stat -c %s $O/r2b_full_k2_de.ncu-rep 2> /dev/null || echo 0)
if [ "$sz" -gt 6000000 ]; then rm -f $O/r2b_full_k2_de.ncu-rep; fi
ls -la $O | grep r2b
