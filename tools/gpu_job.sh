#!/bin/bash
# A/B: downsample convolutions on the lane's side stream vs inline
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -k "executor or lanes or image_equals or golden or bf16_close or full_size" > $OUT/j_tests.log 2>&1
tail -4 $OUT/j_tests.log
for mode in side inline side inline; do
  if [ $mode = inline ]; then export IPSB_DS_INLINE=1; else unset IPSB_DS_INLINE; fi
  for w in traffic mnist; do
    python bench.py --workload $w --steps 20 --skip train,library,cpu,workloads,exact,sustained,seq > $OUT/j_ds.log 2>&1
    python - <<PY
import json
for l in open('$OUT/j_ds.log'):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('$mode $w', round(d['ms_per_step'], 4), round(d['value']), 'busy', round(r['family_busy_ms_per_step'],4), 'frac', round(r['frac'],4))
PY
  done
done
echo "=== done"
