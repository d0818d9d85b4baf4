import sys, os, math, torch
sys.path.insert(0, '/root/repo')
from ips_b200 import ops
dev = torch.device('cuda:0'); P = 1024; H = 25; Cin = Cout = 64
x = ops.to_pf(torch.randn(P, H, H, Cin, device=dev).to(torch.bfloat16))
w = (torch.randn(Cout, 9 * Cin, device=dev) / math.sqrt(9 * Cin)).to(torch.bfloat16)
scale = torch.ones(Cout, device=dev); shift = torch.zeros(Cout, device=dev)
out = torch.zeros((ops.pf_geo(P, H, H)[0], Cout), dtype=torch.bfloat16, device=dev)
for _ in range(8):
    ops.conv_bf16_pf(x, w, scale, shift, None, P, H, H, Cout, 3, 3, 1, 1, True, True, True, out=out)
torch.cuda.synchronize()
