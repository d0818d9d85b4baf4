#!/bin/bash
# scratch job for one gpurun call (edited per call)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:select_loop_cluster --launch-skip 3 --launch-count 1 -o /tmp/sel -f \
    python bench.py --workload camelyon --steps 2 --skip train,library,cpu,workloads,exact,sustained,seq,roofline > $OUT/j_ncu.log 2>&1
ncu -i /tmp/sel.ncu-rep --page details > $OUT/j_sel_details.txt 2>/dev/null
ncu -i /tmp/sel.ncu-rep --page source --csv > $OUT/j_sel_source.csv 2>/dev/null
grep -E "Duration|SM Frequency|Elapsed Cycles" $OUT/j_sel_details.txt | head
timeout 300 ncu --set full --clock-control none --import-source on -k regex:select_loop_cluster --launch-skip 3 --launch-count 1 -o /tmp/sel2 -f \
    python bench.py --workload mnist5000 --steps 2 --skip train,library,cpu,workloads,exact,sustained,seq,roofline > $OUT/j_ncu2.log 2>&1
ncu -i /tmp/sel2.ncu-rep --page details > $OUT/j_sel2_details.txt 2>/dev/null
ncu -i /tmp/sel2.ncu-rep --page source --csv > $OUT/j_sel2_source.csv 2>/dev/null
grep -E "Duration|SM Frequency|Elapsed Cycles" $OUT/j_sel2_details.txt | head
echo "=== done"
