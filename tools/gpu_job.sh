#!/bin/bash
# round-2 final single-GPU session: all tests, smoke, both bench arms, ncu launch list (+DRAM bytes) of the bench command,
# ncu --set full table of one step.  Output -> gpurun_out/f_*.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-300}" "$@" > $OUT/f_$name.log 2>&1; echo "exit $?" | tee -a $OUT/f_$name.log; tail -n "${TAIL:-4}" $OUT/f_$name.log | cut -c1-400; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/f_gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/f_gpu.txt
TAIL=12 TMO=900 run t_gpu python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider
TMO=300 run smoke python __graft_entry__.py smoke
TAIL=1 TMO=600 run bench python bench.py
TAIL=1 TMO=300 run bench_ref python bench.py --impl reference --steps 20 --warmup 5
TAIL=1 TMO=500 run ncu_launches ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 400 --csv --log-file $OUT/f_launches.csv python bench.py --steps 2 --warmup 1 --skip train,library,cpu,workloads,exact,sustained,seq,roofline
TAIL=1 TMO=500 run ncu_full ncu --set full --clock-control none --import-source off \
    -k regex:'stem_pool|conv_pair|conv_halo|conv_tma|stage_s2d|select_loop|gather_rows16' --launch-skip 120 --launch-count 16 \
    -o /tmp/full -f python bench.py --steps 1 --warmup 3 --skip train,library,cpu,workloads,exact,sustained,seq,roofline
ncu -i /tmp/full.ncu-rep --page raw --csv > $OUT/f_full_raw.csv 2>/dev/null
echo "=== done"
