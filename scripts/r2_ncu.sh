#!/bin/bash
# Round 2, profiler call (the only tool used in this gpurun call is ncu).  Every capture follows a plain run of the same
# command that exited 0.  Captures are exported to CSV (raw page) ON THE BOX and only small .ncu-rep files are kept:
# gpurun brings back at most 64 MiB.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/r2_ncu.sh'
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
keep() {  # keep <name>: raw CSV of the capture; the .ncu-rep itself only if it is small
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2> /dev/null
  local sz=$(stat -c %s $O/$1.ncu-rep 2> /dev/null || echo 0)
  if [ "$sz" -gt 6000000 ]; then rm -f $O/$1.ncu-rep; fi
}
# ---- launch lists (cold-cache, serialised: shares only) -------------------------------------------------------------
$B > $O/r2n_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_logncde.csv $B > $O/r2n_l1.log 2>&1
for w in eigenworms_de eigenworms_bd eigenworms_d; do
  $B --workload $w > $O/r2n_plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_$w.csv $B --workload $w > $O/r2n_l_$w.log 2>&1
done
# ---- full captures ---------------------------------------------------------------------------------------------------
python scripts/k1_sweep.py > $O/r2n_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logsig_levy_tma -c 1 -o $O/r2_full_k1_tma python scripts/k1_sweep.py > $O/r2n_c1.log 2>&1
keep r2_full_k1_tma
python scripts/variant_check.py --no-oracle --no-parity --only k2 > $O/r2n_plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:'dlncde_' -s 8 -c 2 -o $O/r2_full_k2d python scripts/variant_check.py --no-oracle --no-parity --only k2 > $O/r2n_c2.log 2>&1
keep r2_full_k2d
$B --workload eigenworms_de > $O/r2n_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'flow_build_tcgen05|de_pack|scan_fwd_cluster|scan_lam_cluster|dlie_t' -s 10 -c 6 \
    -o $O/r2_full_k2_de $B --workload eigenworms_de > $O/r2n_c3.log 2>&1
keep r2_full_k2_de
$B --workload eigenworms_bd > $O/r2n_plain5.log 2>&1 &&
ncu --set full --clock-control none -k regex:'flow_build|pscan_|scan_' -s 12 -c 8 -o $O/r2_full_k2_bd $B --workload eigenworms_bd > $O/r2n_c4.log 2>&1
keep r2_full_k2_bd
$B > $O/r2n_plain6.log 2>&1 &&
ncu --set full --clock-control none -k regex:'logncde_fwd_fast|logncde_bwd_saved|outer_gemm' -s 4 -c 3 -o $O/r2_full_k3 $B > $O/r2n_c5.log 2>&1
keep r2_full_k3
$B --workload toy_logncde > $O/r2n_plain7.log 2>&1 &&
ncu --set full --clock-control none -k regex:'logncde_fwd_kernel|logncde_bwd_kernel' -s 2 -c 2 -o $O/r2_full_k3_toy $B --workload toy_logncde > $O/r2n_c6.log 2>&1
keep r2_full_k3_toy
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --workload eigenworms_nrde > $O/r2n_plain8.log 2>&1 &&
ncu --set full --clock-control none -k regex:'nrde_' -s 2 -c 2 -o $O/r2_full_k3n python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --workload eigenworms_nrde > $O/r2n_c7.log 2>&1
keep r2_full_k3n
du -sm $O; ls -la $O | grep r2_ | head -40
