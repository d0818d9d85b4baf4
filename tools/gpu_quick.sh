#!/bin/bash
# Short GPU session: parity tests + traffic bench (+ optional extra command in $EXTRA).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/$name.log; tail -n "${TAIL:-6}" $OUT/$name.log; }
TAIL=15 TMO=400 run t_kernels python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 120
TAIL=25 TMO=600 run t_ips python -m pytest tests/test_gpu_ips.py -q -m gpu -x --timeout 300
TAIL=3 TMO=400 run bench_traffic_bf16 python bench.py --steps 10 --warmup 3 --no-cpu
TAIL=3 TMO=300 run bench_mnist_bf16 python bench.py --steps 5 --workload mnist --no-cpu
if [ -n "${EXTRA:-}" ]; then TAIL=40 TMO=600 run extra bash -c "$EXTRA"; fi
echo "=== done"
