#!/bin/bash
# scratch job for one gpurun call (edited per call)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -k "select_loop or topm or golden or baseline_size or full_size or lazy or callsite or training_loop" > $OUT/j_tests.log 2>&1
tail -6 $OUT/j_tests.log
for w in mnist5000 mnist traffic; do
    python bench.py --workload $w --steps 10 --skip train,library,cpu,workloads,exact,sustained,seq > $OUT/j_bench_${w}.log 2>&1
    python - <<PY
import json
for l in open('$OUT/j_bench_${w}.log'):
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['ms_per_step'], 4), d['value'])
PY
done
python tools/select_sweep.py > $OUT/j_select_sweep.txt 2>&1; tail -30 $OUT/j_select_sweep.txt
echo "=== done"
