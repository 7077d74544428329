#!/bin/bash
# First GPU call of round 2 (one gpurun, ~6-8 GPU-minutes): everything that was written blind at the end of
# round 1 gets its hardware verdict, and the captures that profiles/README.md lists as gaps are taken.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/round2_gpu_plan.sh'
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
# one process per family (k1 k3 k2 k2de k3g k2p): a fault in one does not lose the others
for fam in k1 k3 k2 k2de k3g k2p; do
  timeout 300 python scripts/variant_check.py --only $fam >> gpurun_out/r2_variants.json 2>> gpurun_out/r2_variants.err
done
timeout 300 python scripts/phase_timing.py > gpurun_out/r2_phases.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
for w in eigenworms_bd eigenworms_d eigenworms_de scp1_bd eigenworms_nrde toy_logncde ppg_logncde_d1; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline >> gpurun_out/r2_bench_others.json 2>> gpurun_out/r2_bench.err
done
# BASELINE configs[4]: DE vs D LNCDE on depth-3 rows, batch / length sweep (quick grid; the full grid and N > 1 later)
timeout 600 python scripts/lncde_sweep.py --quick --out gpurun_out/r2_lncde_sweep.json > gpurun_out/r2_lncde_sweep.log 2>&1
# the same points with the variants that target these shapes: fused diagonal kernels (HBM path of the D-LNCDE), tcgen05
# flow build / bracket gradient for the dense model (depth-3 rows: n_basis = 140)
LNCDE_TUNE_K2_FUSED_D=1 LNCDE_TUNE_K2_DE_V2=3 timeout 600 python scripts/lncde_sweep.py --quick \
    --out gpurun_out/r2_lncde_sweep_variants.json > gpurun_out/r2_lncde_sweep_variants.log 2>&1
python scripts/k1_sweep.py > gpurun_out/r2_k1_sweep.txt 2>&1
LNCDE_TUNE_K1_TMA=1 python scripts/k1_sweep.py > gpurun_out/r2_k1_sweep_tma.txt 2>&1
# launch list of the default step (shares only), then one full capture of the K1 TMA kernel and of the K2 kernels
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune --no-graph > gpurun_out/r2_ncu1.log 2>&1
LNCDE_TUNE_K1_TMA=1 python scripts/k1_sweep.py > gpurun_out/r2_plain2.log 2>&1 &&
LNCDE_TUNE_K1_TMA=1 ncu --set full --clock-control none --import-source on -k regex:logsig_levy -c 2 \
    -o gpurun_out/r2_prof_k1_tma python scripts/k1_sweep.py > gpurun_out/r2_ncu2.log 2>&1
python bench.py --workload eigenworms_de --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'scan_|flow_build' -s 6 -c 4 \
    -o gpurun_out/r2_prof_k2 python bench.py --workload eigenworms_de --steps 2 --warmup 3 --no-cpu-baseline --no-graph \
    > gpurun_out/r2_ncu3.log 2>&1
# tensor-pipe variant of the dense LNCDE (k2_de_v2 = 2): one capture of the 3xTF32 flow build and bracket gradient
LNCDE_TUNE_K2_DE_V2=2 python bench.py --workload eigenworms_de --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-autotune \
    > gpurun_out/r2_plain4.log 2>&1 &&
LNCDE_TUNE_K2_DE_V2=2 ncu --set full --clock-control none --import-source on -k regex:'flow_build_tc|dlie_triple_tc|scan_bwd_tma_lam' \
    -s 6 -c 3 -o gpurun_out/r2_prof_k2_tc python bench.py --workload eigenworms_de --steps 2 --warmup 3 --no-cpu-baseline \
    --no-graph --no-autotune > gpurun_out/r2_ncu4.log 2>&1
# tcgen05 variant of the dense flow build (k2_de_v2 = 3): bench line, then one capture of the CUTLASS FastFP32 kernel + packing
LNCDE_TUNE_K2_DE_V2=3 python bench.py --workload eigenworms_de --steps 5 --warmup 3 --no-cpu-baseline --no-graph --no-autotune \
    > gpurun_out/r2_de_umma.json 2> gpurun_out/r2_de_umma.err &&
LNCDE_TUNE_K2_DE_V2=3 ncu --set full --clock-control none --import-source on -k regex:'cutlass|umma_pack' -s 6 -c 3 \
    -o gpurun_out/r2_prof_k2_umma python bench.py --workload eigenworms_de --steps 2 --warmup 3 --no-cpu-baseline \
    --no-graph --no-autotune > gpurun_out/r2_ncu5.log 2>&1
tail -3 gpurun_out/r2_pytest.log; head -c 1500 gpurun_out/r2_variants.json; echo; head -c 600 gpurun_out/r2_bench.json
