#!/bin/bash
# One GPU session (about 6 GPU-minutes): parity tests, smoke, benches of the four workloads, the ncu launch list with DRAM
# traffic that profiles/r01_traffic.json is derived from, and an ncu --set full table of one steady-state step exported
# as CSV on the box (the .ncu-rep itself stays in /tmp: gpurun copies back at most 64 MiB).  Output -> gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
#   python tools/launch_traffic.py gpurun_out/launches.csv profiles/rNN_traffic.json
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/$name.log; tail -n "${TAIL:-4}" $OUT/$name.log | cut -c1-400; }
TAIL=6 TMO=600 run t_gpu python -m pytest tests -q -m gpu -x --timeout 300
TMO=300 run smoke python __graft_entry__.py smoke
TAIL=1 TMO=400 run bench_traffic python bench.py
for w in mnist mnist5000 camelyon; do TAIL=1 TMO=300 run bench_$w python bench.py --workload $w --steps 10 --warmup 3 --no-cpu; done
TAIL=20 TMO=300 run conv_bench python tools/conv_bench.py 1024 all
TAIL=1 TMO=500 run ncu_launches ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 700 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-train
TAIL=1 TMO=500 run ncu_full ncu --set full --clock-control none --import-source off \
    -k regex:'stem_pool|conv_pair|conv_halo|conv_tma|stage_s2d|select_loop|gather_rows16' --launch-skip 120 --launch-count 16 \
    -o /tmp/full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-train
ncu -i /tmp/full.ncu-rep --page raw --csv > $OUT/full_raw.csv 2>/dev/null
echo "=== done"
