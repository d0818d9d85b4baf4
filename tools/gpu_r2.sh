#!/bin/bash
# Round-2 GPU session: parity tests (all, not fail-fast), smoke, both bench arms.  Output -> gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_r2.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
TAG=${TAG:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/${TAG}_gpu.txt
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/${TAG}_$name.log 2>&1; echo "exit $?" | tee -a $OUT/${TAG}_$name.log; tail -n "${TAIL:-4}" $OUT/${TAG}_$name.log | cut -c1-600; }
if [ -z "${NOTEST:-}" ]; then TAIL=25 TMO=900 run t_gpu python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider ${PYTEST_ARGS:-}; fi
if [ -z "${NOSMOKE:-}" ]; then TMO=300 run smoke python __graft_entry__.py smoke; fi
if [ -z "${NOBENCH:-}" ]; then
  TAIL=1 TMO=600 run bench python bench.py ${BENCH_ARGS:-}
  TAIL=1 TMO=300 run bench_ref python bench.py --impl reference --steps 5 --warmup 1
fi
if [ -n "${EXTRA:-}" ]; then TAIL=40 TMO=${EXTRA_TMO:-600} run extra bash -c "$EXTRA"; fi
echo "=== done"
