#!/bin/bash
# scratch job: parallel BN reductions + TMA staging: tests, bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/s_$name.log 2>&1; echo "exit $?" | tee -a $OUT/s_$name.log; tail -n "${TAIL:-4}" $OUT/s_$name.log | cut -c1-600; }
TAIL=25 TMO=900 run t_all python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider
TAIL=1 TMO=400 run bench python bench.py --skip library,cpu,exact,sustained,workloads,seq
TAIL=1 TMO=400 run bench_notma env IPSB_STAGE_NO_TMA=1 python bench.py --skip library,cpu,exact,sustained,workloads,seq,train
echo "=== done"
