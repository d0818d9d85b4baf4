"""e2e (host tensor -> ips() -> winners on the host) vs stream-in chunk size, with the raw H2D time as the floor."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ips_b200 import IPSNet

dev = torch.device('cuda:0')
conf, B, N = bench.conf_for('traffic', 'bf16')
torch.manual_seed(0)
x = torch.randn(B, N, 3, 100, 100)
xh = x.pin_memory()
d = torch.empty_like(x, device=dev)


def timed(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timed(lambda: d.copy_(xh, non_blocking=True))
print(f'raw H2D {x.numel()*4/1e6:.0f} MB: {ms:.2f} ms = {x.numel()*4/ms/1e6:.1f} GB/s')
res_h = torch.empty((B, conf.M, 3, 100, 100)).pin_memory()
for chunk in [256, 384, 512, 768, 1024, 1536]:
    os.environ['IPS_B200_CHUNK'] = str(chunk)
    net = IPSNet(dev, conf).to(dev).eval()
    def step():
        mp, _ = net.ips(xh)
        res_h.copy_(mp, non_blocking=True)
    step(); step()
    ms = timed(step)
    print(f'chunk {chunk:5d}: {ms:.2f} ms  {B*N/ms*1e3/1e3:.0f} k patches/s')
