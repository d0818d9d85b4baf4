#!/bin/bash
# Short GPU session (about 1.5 GPU-minutes): parity tests + traffic / mnist benches (+ optional extra command in $EXTRA).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/$name.log; tail -n "${TAIL:-4}" $OUT/$name.log | cut -c1-300; }
TAIL=5 TMO=600 run t_gpu python -m pytest tests -q -m gpu -x --timeout 300
TAIL=1 TMO=300 run bench_traffic python bench.py --steps 10 --warmup 3 --no-cpu --no-train
TAIL=1 TMO=300 run bench_mnist python bench.py --steps 10 --warmup 3 --workload mnist --no-cpu --no-train
if [ -n "${EXTRA:-}" ]; then TAIL=40 TMO=600 run extra bash -c "$EXTRA"; fi
echo "=== done"
