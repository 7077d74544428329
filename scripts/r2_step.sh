#!/bin/bash
# Round 2, targeted check of one change: the named GPU test files, the named variant families, two bench lines.
#   scripts/r2_step.sh "tests/test_linear_ncde_gpu.py tests/test_logncde_gpu.py" "k2de k3" "eigenworms_de"
set -u
O=gpurun_out
mkdir -p $O
TESTS=${1:-tests}
FAMS=${2:-}
WLS=${3:-}
timeout -k 10 1200 python -m pytest $TESTS -m gpu -q -x > $O/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2s_pytest.log; tail -5 $O/r2s_pytest.log
for f in $FAMS; do
  timeout -k 10 400 python scripts/variant_check.py --no-oracle --only $f > $O/r2s_variant_$f.json 2> $O/r2s_variant_$f.err
  echo "variant $f rc=$?"; head -c 1800 $O/r2s_variant_$f.json; echo
done
timeout -k 10 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2s_bench.json 2> $O/r2s_bench.err; head -c 400 $O/r2s_bench.json; echo
rm -f $O/r2s_bench_others.json
for w in $WLS; do
  timeout -k 10 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline >> $O/r2s_bench_others.json 2>> $O/r2s_bench.err
done
[ -f $O/r2s_bench_others.json ] && python - <<'PY'
import json
for l in open('gpurun_out/r2s_bench_others.json'):
    d = json.loads(l)
    print(d['metric'][22:], '%.3f ms/step' % d['ms_per_step'], '%.0f samples/s' % d['value'])
PY
