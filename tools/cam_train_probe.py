import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from ips_b200 import IPSNet
from ips_b200.train import GraphedTrainStep
dev = torch.device('cuda:0')
conf, B, N = bench.conf_for('camelyon', 'bf16')
net = IPSNet(dev, conf).to(dev).train()
x = torch.randn(B, N, conf.n_chan_in, device=dev)
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=conf.wd, capturable=True)
g = GraphedTrainStep(net, conf, opt, 16)
def timed(f, n=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print('ips', timed(lambda: net.ips(x)))
print('ips out', timed(lambda: net.ips(x, out=g.buffers, row_offset=3)))
for i in range(16): net.ips(x, out=g.buffers, row_offset=i)
g.capture()
print('graph step', timed(lambda: g()))
print('eager step', timed(lambda: g._step(), 3))
