#!/bin/bash
# Round-2 multi-GPU session (N = $NG GPUs of one box): 2-rank NCCL / peer-memory tests, then the bench under torchrun.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'NG=2 bash tools/gpu_r2_multi.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
NG=${NG:-2}
TAG=${TAG:-r2_n$NG}
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/${TAG}_$name.log 2>&1; echo "exit $?" | tee -a $OUT/${TAG}_$name.log; tail -n "${TAIL:-4}" $OUT/${TAG}_$name.log | cut -c1-600; }
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
if [ -z "${NOTEST:-}" ]; then TAIL=30 TMO=600 run t_multi python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 300 -p no:cacheprovider ${K_EXPR:+-k "$K_EXPR"}; fi
if [ -z "${NOBENCH:-}" ]; then
  TAIL=3 TMO=${BENCH_TMO:-800} run bench python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $NG --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS:-}
fi
if [ -n "${EXTRA:-}" ]; then TAIL=40 TMO=${EXTRA_TMO:-600} run extra bash -c "$EXTRA"; fi
echo "=== done"
