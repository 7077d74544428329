#!/bin/bash
# Round 2, measurement call (no profiler): parity errors of every GPU test, default bench line, every kernel variant with
# its timing, per-workload bench lines, phase timing, BASELINE configs[4] quick sweep.  Text outputs only (gpurun brings
# back at most 64 MiB).
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/r2_measure.sh'
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r2_smoke.log 2>&1
LNCDE_ERR_LOG=$PWD/$O/r2_parity_errors.jsonl timeout 1500 python -m pytest tests -m gpu -q > $O/r2_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2_pytest.log
python bench.py --steps 20 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err
for fam in k1 k3 k2 k2de k3g k2p; do
  timeout 300 python scripts/variant_check.py --no-oracle --only $fam >> $O/r2_variants.json 2>> $O/r2_variants.err
done
timeout 120 python scripts/variant_check.py --no-oracle --only k2de --levels 3 >> $O/r2_variants_tc5.json 2>> $O/r2_variants.err
timeout 300 python scripts/phase_timing.py > $O/r2_phases.txt 2>&1
for w in eigenworms_bd eigenworms_d eigenworms_de scp1_bd eigenworms_nrde toy_logncde ppg_logncde_d1 ppg_logncde_d2; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline >> $O/r2_bench_others.json 2>> $O/r2_bench.err
done
timeout 600 python scripts/lncde_sweep.py --quick --out $O/r2_lncde_sweep.json > $O/r2_lncde_sweep.log 2>&1
LNCDE_TUNE_K2_FUSED_D=1 LNCDE_TUNE_K2_DE_V2=2 timeout 600 python scripts/lncde_sweep.py --quick \
    --out $O/r2_lncde_sweep_variants.json > $O/r2_lncde_sweep_variants.log 2>&1
python scripts/k1_sweep.py > $O/r2_k1_sweep.txt 2>&1
LNCDE_TUNE_K1_TMA=1 python scripts/k1_sweep.py > $O/r2_k1_sweep_tma.txt 2>&1
du -sm $O; tail -3 $O/r2_pytest.log; head -c 400 $O/r2_bench.json; echo; cat $O/r2_variants_tc5.json | head -c 1500
