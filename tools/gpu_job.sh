#!/bin/bash
# scratch job: gradients created by backward (no zero fill), torch's fused AdamW in the bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/s_$name.log 2>&1; echo "exit $?" | tee -a $OUT/s_$name.log; tail -n "${TAIL:-4}" $OUT/s_$name.log | cut -c1-600; }
TAIL=15 TMO=800 run t_train python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -k "train or step or graph or callsite"
TAIL=1 TMO=400 run bench python bench.py --skip library,cpu,exact,sustained,workloads,seq,staging,roofline
echo "=== done"
