#!/bin/bash
# bf16x3 precision: parity report + per-call timing
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -s -k "bf16x3" 2>&1 | grep -E "bf16x3|passed|failed" > $OUT/j_tests.log
cat $OUT/j_tests.log | cut -c1-220
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from ips_b200 import ops
class A: pass
b = bench.Bench(type('X', (), {'skip': '', 'features': 'fp32'})())
m = b.measure('traffic', 'bf16x3', 3, 2, e2e=False)
ops.TIMER = {}
m['net'].ips(m['x']); torch.cuda.synchronize()
per = {k: (sum(a.elapsed_time(bb) for a, bb, _ in v), len(v)) for k, v in ops.TIMER.items()}
ops.TIMER = None
for k, (ms, n) in sorted(per.items(), key=lambda kv: -kv[1][0]): print(f'{k:36s} {ms:8.3f} ms  {n:4d} launches')
for w in ('mnist', 'camelyon'):
    mm = b.measure(w, 'bf16x3', 3, 2, e2e=False)
    print(w, 'bf16x3', round(mm['value']), round(mm['ms_per_step'], 3))
PY
echo "=== done"
