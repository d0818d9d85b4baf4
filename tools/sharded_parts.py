"""Single-GPU timing of the kernels one rank runs in the sequence-sharded CAMELYON call (R = 8, one 200 k / 50 k bag):
encoder on the slice, local selection loop, candidate merge (scores + stable top-M on R*M entries), final gather."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ips_b200 import IPSNet, Struct, ops
from ips_b200.configs import load_config

dev = torch.device('cuda:0')
conf = load_config('camelyon', precision='bf16')
net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
with torch.no_grad():
    net.transf.crs_attn.q.mul_(12.0)
ca = net.transf.crs_attn
M, I, R = conf.M, conf.I, 8


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for B, n_local in ((1, 25000), (1, 6250), (16, 6250)):
    x = torch.randn(B, n_local, 2048, device=dev)
    z = net.patch_logits(x)
    L = R * M
    cz = torch.randn(B, L, 8, device=dev)
    sc = ops.scores_from_logits(cz, ca.H, ca.n_token)
    win = ops.topm_stable(sc, M)[1]
    print(f'B={B} n_local={n_local}: encoder {t(lambda: net.patch_logits(x)):8.1f} us | local loop {t(lambda: ops.select_loop(z, None, False, ca.H, ca.n_token, M, I)):8.1f} us | '
          f'scores(L={L}) {t(lambda: ops.scores_from_logits(cz, ca.H, ca.n_token)):8.1f} us | topm(L={L}) {t(lambda: ops.topm_stable(sc, M)):8.1f} us | '
          f'gather {t(lambda: ops.gather_rows(x, win % n_local, n_local)):8.1f} us')
