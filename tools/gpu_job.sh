#!/bin/bash
# scratch job: 3-pass radix select in the cluster loop: parity tests + timings
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/s_$name.log 2>&1; echo "exit $?" | tee -a $OUT/s_$name.log; tail -n "${TAIL:-4}" $OUT/s_$name.log | cut -c1-600; }
TAIL=15 TMO=800 run t_sel python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ips.py -q -m gpu --timeout 300 -p no:cacheprovider -k "select or streamed or merge or ips or topm or golden"
TAIL=14 TMO=200 run timing python tools/streamed_timing.py
TAIL=1 TMO=300 run bench_m5000 python bench.py --workload mnist5000 --skip library,cpu,exact,sustained,workloads,seq,train,staging
echo "=== done"
