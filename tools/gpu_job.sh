#!/bin/bash
# scratch job: streamed selection + device scan order: tests, timing, bench records
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/s_$name.log 2>&1; echo "exit $?" | tee -a $OUT/s_$name.log; tail -n "${TAIL:-4}" $OUT/s_$name.log | cut -c1-600; }
TAIL=15 TMO=600 run t_new python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -k "streamed or select_loop or projector or merge or keyed"
TAIL=14 TMO=200 run timing python tools/streamed_timing.py
TAIL=1 TMO=400 run bench python bench.py --skip train,library,cpu,exact,sustained
echo "=== done"
