#!/bin/bash
# scratch job for one gpurun call (edited per call)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -k "projector_logits_fused or camelyon or bf16_feature_bag" > $OUT/j_tests.log 2>&1
tail -3 $OUT/j_tests.log
for pf in 1 0; do
    IPSB_PROJ_PREFETCH=$pf python bench.py --workload camelyon --steps 20 --skip train,library,cpu,workloads,exact,sustained,seq > $OUT/j_bench_pf$pf.log 2>&1
    python - <<PY
import json
for l in open('$OUT/j_bench_pf$pf.log'):
    if l.startswith('{'):
        d = json.loads(l); print('prefetch $pf', round(d['ms_per_step'], 4), d['roofline']['kernel_ms_all'] if d.get('roofline') else None)
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:projector_logits --launch-skip 3 --launch-count 1 -o /tmp/proj -f \
    python bench.py --workload camelyon --steps 2 --skip train,library,cpu,workloads,exact,sustained,seq,roofline > $OUT/j_ncu.log 2>&1
ncu -i /tmp/proj.ncu-rep --page details > $OUT/j_proj_details.txt 2>/dev/null
ncu -i /tmp/proj.ncu-rep --page source --csv > $OUT/j_proj_source.csv 2>/dev/null
grep -E "Duration|SM Frequency|DRAM Throughput|L2 Cache Throughput|TC is" $OUT/j_proj_details.txt | head
echo "=== done"
