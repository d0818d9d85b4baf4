#!/bin/bash
# scratch job: pinned scan-order staging on traffic: A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/s_$name.log 2>&1; echo "exit $?" | tee -a $OUT/s_$name.log; tail -n "${TAIL:-4}" $OUT/s_$name.log | cut -c1-300; }
for rep in 1 2; do
for v in 1 0; do
TAIL=1 TMO=200 run pin_${v}_$rep env IPS_B200_PINNED_ORDER=$v python bench.py --skip library,cpu,exact,sustained,workloads,seq,train,roofline,staging
done
done
TAIL=1 TMO=200 run pin_mnist_1 env IPS_B200_PINNED_ORDER=1 python bench.py --workload mnist --skip library,cpu,exact,sustained,workloads,seq,train,roofline,staging
TAIL=1 TMO=200 run pin_mnist_0 env IPS_B200_PINNED_ORDER=0 python bench.py --workload mnist --skip library,cpu,exact,sustained,workloads,seq,train,roofline,staging
echo "=== done"
