"""Per-layer timing of the tcgen05 conv kernel on the traffic / mnist encoder shapes."""
import sys, os, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ips_b200 import ops

dev = torch.device('cuda:0')
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != 'all' else None
MODE0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
# name, Cin, Cout, k, stride, pad, H
LAYERS = [
    ('stem', 4, 64, 7, 2, 3, 100),
    ('l1', 64, 64, 3, 1, 1, 25),
    ('l2s', 64, 128, 3, 2, 1, 25),
    ('l2d', 64, 128, 1, 2, 0, 25),
    ('l2', 128, 128, 3, 1, 1, 13),
    ('l3s', 128, 256, 3, 2, 1, 13),
    ('l3', 256, 256, 3, 1, 1, 7),
    ('l4s', 256, 512, 3, 2, 1, 7),
    ('l4', 512, 512, 3, 1, 1, 4),
]
tot_ms = 0
for name, Cin, Cout, k, s, p, H in LAYERS:
    if only and name != only:
        continue
    stem3 = (Cin == 4 and MODE0 == 0)
    Hx = H + 6 if stem3 else H
    x = torch.randn(P, Hx, Hx, Cin, device=dev).to(torch.bfloat16)
    Kp = 256 if Cin == 4 else k * k * Cin
    w = (torch.randn(Cout, Kp, device=dev) / math.sqrt(Kp)).to(torch.bfloat16)
    scale = torch.ones(Cout, device=dev); shift = torch.zeros(Cout, device=dev)
    mode = (3 if stem3 else 1) if Cin == 4 else MODE0
    f = lambda: ops.conv_bf16(x, w, scale, shift, None, Cout, k, k, s, p, True, mode)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    Ho = (H + 2 * p - k) // s + 1
    kreal = (147 if Cin == 4 else k * k * Cin)
    fl = 2.0 * P * Ho * Ho * Cout * kreal
    print(f'{name:5s} M={P*Ho*Ho:7d} N={Cout:3d} K={Kp:4d}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (algorithmic)  tiles={-(-P*Ho*Ho//128)}x{Cout//(128 if Cout%128==0 else 64)}')

# ---- padded-flat shifted-window kernel
print('--- PF shifted-window 3x3 (halo) kernel')
for name, Cin, Cout, H in [('l1', 64, 64, 25), ('l2', 128, 128, 13), ('m_l1', 64, 64, 13), ('m_l2', 128, 128, 7)]:
    if only and name != only:
        continue
    x = ops.to_pf(torch.randn(P, H, H, Cin, device=dev).to(torch.bfloat16))
    w = (torch.randn(Cout, 9 * Cin, device=dev) / math.sqrt(9 * Cin)).to(torch.bfloat16)
    scale = torch.ones(Cout, device=dev); shift = torch.zeros(Cout, device=dev)
    out = torch.zeros((ops.pf_geo(P, H, H)[0], Cout), dtype=torch.bfloat16, device=dev)
    f = lambda: ops.conv_bf16_pf(x, w, scale, shift, None, P, H, H, Cout, 3, 3, 1, 1, True, True, True, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 2.0 * P * H * H * Cout * 9 * Cin
    print(f'{name:5s} P={P} {H}x{H} {Cin}->{Cout}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (algorithmic)')

# ---- space-to-depth stem chain: staging, shifted-window stem, strided max-pool
print('--- s2d stem chain')
if not only or only == 's2d':
    for C, H in [(3, 100), (1, 50)]:
        x = torch.randn(P, C, H, H, device=dev)
        w = (torch.randn(64, 256, device=dev) / 16).to(torch.bfloat16)
        scale = torch.ones(64, device=dev); shift = torch.zeros(64, device=dev)
        Ho, Wp = H // 2, H // 2 + 3
        Sp = (Ho + 3) * Wp
        frame = ops.stage_patches_s2d(x, P, C, H, H)
        y = ops.conv_stem_s2d(frame, w, scale, shift, P, H, H)
        pooled = ops.maxpool3x3s2_pf_strided(y, P, Ho, Ho, 64, Wp, Sp)
        for name, f in [('stage', lambda: ops.stage_patches_s2d(x, P, C, H, H)),
                        ('stem', lambda: ops.conv_stem_s2d(frame, w, scale, shift, P, H, H)),
                        ('pool', lambda: ops.maxpool3x3s2_pf_strided(y, P, Ho, Ho, 64, Wp, Sp, out=pooled)),
                        ('fused', lambda: ops.stem_pool_s2d(frame, w, scale, shift, P, H, H, out=pooled))]:
            for _ in range(3): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize()
            print(f'{name:5s} C={C} {H}x{H} P={P}  {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us')
