#!/bin/bash
# scratch job for one gpurun call (edited per call)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -k "select_loop or topm or merge_candidates or camelyon" > $OUT/j_tests.log 2>&1
tail -8 $OUT/j_tests.log
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from ips_b200 import ops
dev = torch.device('cuda:0')
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for B, L in ((1, 40000), (16, 40000), (1, 20000), (1, 10000)):
    cz = torch.randn(B, L, 8, device=dev)
    print(f'merge B={B} L={L}: cluster loop {t(lambda: ops.merge_candidates(cz, 8, 1, 5000)):7.1f} us | scores+topm {t(lambda: ops.topm_stable(ops.scores_from_logits(cz, 8, 1), 5000)):7.1f} us')
PY
echo "=== done"
