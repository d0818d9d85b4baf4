import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ips_b200 import ops
dev = torch.device('cuda:0')
def run(P, H, Cin, Cout, tag):
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
    tiles = (P * (H + 1) * (H + 1) + 127) // 128
    print(f'{tag} dbg={os.environ.get("IPSB_DEBUG","0")} P={P} {H}x{H} {Cin}->{Cout}: {ms*1e3:7.1f} us  {tiles/148:.1f} tiles/SM  {ms*1e3/(tiles/148):.2f} us/tile')
for P in (1024, 128):
    run(P, 25, 64, 64, 'l1')
run(1024, 13, 128, 128, 'l2')
