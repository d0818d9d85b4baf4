"""One eager train step (forward + loss + backward + AdamW on the winners of one traffic batch) for an ncu launch list."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ips_b200 import IPSNet
from ips_b200.train import GraphedTrainStep

dev = torch.device('cuda:0')
wl = sys.argv[1] if len(sys.argv) > 1 else 'traffic'
conf, B, N = bench.conf_for(wl, 'bf16')
net = IPSNet(dev, conf).to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=conf.wd, capturable=True)
step = GraphedTrainStep(net, conf, opt, B)
step.mem_patch.normal_()
for _ in range(3):
    step._step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step._step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
