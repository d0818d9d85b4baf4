#!/bin/bash
# ncu --set full of the projector kernel (final state) -> gpurun_out/p_proj_details.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:projector_logits --launch-skip 3 --launch-count 1 -o /tmp/proj -f \
    python bench.py --workload camelyon --steps 2 --skip train,library,cpu,workloads,exact,sustained,seq,roofline > $OUT/p_ncu.log 2>&1
ncu -i /tmp/proj.ncu-rep --page details > $OUT/p_proj_details.txt 2>/dev/null
grep -E "Duration|SM Frequency|DRAM Throughput|L2 Cache Throughput|TC is|SM Active Cycles|Elapsed Cycles" $OUT/p_proj_details.txt | head
echo "=== done"
