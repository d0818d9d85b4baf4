"""Selection microbench sweep (BASELINE.json configs[4]): embeddings (B, N, 512) -> logits ->
sequential top-M loop -> gather of the winners; N 1k..200k x M 10..5000, I = M (and the traffic-style M=10, I=32),
n_token T in {1, 4} (H = 8).  Prints one line per point: time of (logits + select + gather), patches/s, the HBM
fraction of the algorithmic bytes (N*D*4 read once + 2*M*D*4 for the gather) and, for N in {10 000, 50 000} with
`--cpu`, the reference's loop (oracle `score_and_select` iterated, ips_net.py:136-155,218-241) on the host cores.

    python tools/select_sweep.py [B] [--cpu]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ips_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
D, H = 512, 8
CPU = '--cpu' in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith('--')]
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
except Exception:
    pass
HBM = peaks.get('hbm_gbs', 6650.0)
B = int(args[0]) if args else 1
torch.manual_seed(0)
rows = []


def cpu_loop_ms(emb_cpu, perm_cpu, M, I, T):
    """The reference loop on the host: cat(memory, next I) -> get_scores -> topk, with the oracle's functions."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import math
    import ips_oracle as O
    conf = O.preset('camelyon', M=M, I=I, H=H, n_token=T, D=D, D_k=64, D_v=64, n_chan_in=16, shuffle=False)
    conf.tasks = {f't{t}': {'id': t, 'name': f't{t}', 'act_fn': 'softmax', 'metric': 'accuracy'} for t in range(T)}
    sd = O.make_state(conf, 1, q_gain=12.0)
    torch.set_num_threads(os.cpu_count())
    Bc, N = emb_cpu.shape[:2]
    t0 = time.perf_counter()
    with torch.no_grad():
        px = emb_cpu[:, perm_cpu[0]]
        idx = torch.arange(N).unsqueeze(0).expand(Bc, -1)
        mem_emb, mem_idx = px[:, :M], idx[:, :M]
        for it in range(math.ceil((N - M) / I)):
            lo = M + it * I
            hi = min(lo + I, N)
            mem_emb, mem_idx, _ = O.score_and_select(sd, conf, torch.cat((mem_emb, px[:, lo:hi]), 1), None, M,
                                                     torch.cat((mem_idx, idx[:, lo:hi]), 1))
    return 1e3 * (time.perf_counter() - t0)


for T in (1, 4):
  U = torch.randn(D, H * T, device=dev) * 0.2
  for N in (1000, 2000, 5000, 10000, 20000, 50000, 100000, 200000):
    emb = torch.randn(B, N, D, device=dev)
    perm = torch.randperm(N, device=dev).unsqueeze(0)
    emb_cpu = emb.cpu() if CPU and N in (10000, 50000) else None
    for M, I in ((10, 32), (10, 10), (50, 50), (100, 100), (500, 500), (1000, 1000), (5000, 5000)):
        if M >= N or (T == 4 and (M, I) in ((10, 10), (50, 50), (500, 500))):
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
        cpu = ''
        if emb_cpu is not None and n_iter <= 5000:
            cms = cpu_loop_ms(emb_cpu, perm.cpu(), M, I, T)
            cpu = f'  | CPU reference loop {cms:9.1f} ms ({os.cpu_count()} threads) = {cms / ms:7.0f}x'
        rows.append((N, M, I, n_iter, ms, B * N / ms * 1e3, bytes_alg / ms / 1e6 / HBM))
        print(f'B={B:2d} T={T} N={N:6d} M={M:4d} I={I:4d} iters={n_iter:5d}  {ms*1e3:9.1f} us  {ms*1e3/n_iter:7.2f} us/iter  '
              f'{B*N/ms*1e3/1e6:8.2f} Mpatch/s  {bytes_alg/ms/1e6:7.1f} GB/s = {bytes_alg/ms/1e6/HBM:5.3f} of HBM peak{cpu}', flush=True)
