#!/bin/bash
# Round 2, single-GPU check after a change: smoke, GPU tests, default bench line, the other workloads, quick sweep.
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r2c_smoke.log 2>&1; tail -1 $O/r2c_smoke.log
LNCDE_ERR_LOG=$PWD/$O/r2c_parity_errors.jsonl timeout 1500 python -m pytest tests -m gpu -q > $O/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest.log; tail -4 $O/r2c_pytest.log
python bench.py --steps 20 --warmup 3 > $O/r2c_bench.json 2> $O/r2c_bench.err; head -c 300 $O/r2c_bench.json; echo
rm -f $O/r2c_bench_others.json
for w in eigenworms_bd eigenworms_d eigenworms_de scp1_bd eigenworms_nrde toy_logncde ppg_logncde_d1 ppg_logncde_d2 eigenworms_logncde_c6; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline >> $O/r2c_bench_others.json 2>> $O/r2c_bench.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c_bench_others.json'):
    d = json.loads(l)
    print(d['metric'][22:], '%.3f ms/step' % d['ms_per_step'], '%.0f samples/s' % d['value'])
PY
timeout 600 python scripts/lncde_sweep.py --quick --out $O/r2c_lncde_sweep.json > $O/r2c_lncde_sweep.log 2>&1; tail -1 $O/r2c_lncde_sweep.log | cut -c1-500
timeout 300 python scripts/variant_check.py --no-oracle --no-parity > $O/r2c_variants_timing.json 2> $O/r2c_variants.err; head -c 2500 $O/r2c_variants_timing.json
