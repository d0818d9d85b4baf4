#!/bin/bash
# One GPU session: unit + e2e parity tests (separate processes so a trap in one file does
# not poison the others), smoke, benches, and an ncu launch list.  Output -> gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/$name.log; tail -n "${TAIL:-6}" $OUT/$name.log; }
TMO=240 run probe python tools/umma_probe.py
TAIL=25 TMO=400 run t_kernels_simt python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not umma" -x --timeout 120
TAIL=25 TMO=400 run t_kernels_umma python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "umma" --timeout 120
TAIL=30 TMO=600 run t_ips python -m pytest tests/test_gpu_ips.py -q -m gpu -s --timeout 300
TMO=300 run smoke python __graft_entry__.py smoke
TAIL=3 TMO=400 run bench_traffic_bf16 python bench.py --steps 5 --warmup 3
TAIL=3 TMO=400 run bench_traffic_fp32 python bench.py --steps 2 --warmup 3 --precision fp32 --no-cpu
TAIL=3 TMO=300 run bench_camelyon_bf16 python bench.py --steps 10 --workload camelyon --no-cpu
TAIL=3 TMO=300 run bench_mnist_bf16 python bench.py --steps 5 --workload mnist --no-cpu
TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu
echo "=== done"
