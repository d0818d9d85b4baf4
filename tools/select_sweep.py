"""Selection microbench sweep (BASELINE.json configs[4]): embeddings (B, N, 512) -> logits ->
sequential top-M loop -> gather of the winners; N 1k..200k x M 10..5000, I = M (and the traffic-style M=10, I=32).
Prints one line per point: time of (logits + select + gather), patches/s, and the HBM fraction of the
algorithmic bytes (N*D*4 read once + 2*M*D*4 for the gather)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ips_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
D, H, T = 512, 8, 1
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
except Exception:
    pass
HBM = peaks.get('hbm_gbs', 6650.0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
U = torch.randn(D, H * T, device=dev) * 0.2
rows = []
for N in (1000, 2000, 5000, 10000, 20000, 50000, 100000, 200000):
    emb = torch.randn(B, N, D, device=dev)
    perm = torch.randperm(N, device=dev).unsqueeze(0)
    for M, I in ((10, 32), (10, 10), (50, 50), (100, 100), (500, 500), (1000, 1000), (5000, 5000)):
        if M >= N:
            continue

        def step():
            z = ops.logits(emb.view(B * N, D), U).view(B, N, H * T)
            _, src, _ = ops.select_loop(z, perm, False, H, T, M, I)
            return ops.gather_rows(emb, src, N)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_alg = B * (N * D * 4 + 2 * M * D * 4)
        n_iter = -(-(N - M) // I)
        rows.append((N, M, I, n_iter, ms, B * N / ms * 1e3, bytes_alg / ms / 1e6 / HBM))
        print(f'N={N:6d} M={M:4d} I={I:4d} iters={n_iter:5d}  {ms*1e3:9.1f} us  {B*N/ms*1e3/1e6:8.2f} Mpatch/s  '
              f'{bytes_alg/ms/1e6:7.1f} GB/s = {bytes_alg/ms/1e6/HBM:5.3f} of HBM peak', flush=True)
