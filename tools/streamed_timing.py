"""Feature-bag ips() with the projector and the selection loop side by side vs one after the other (ms per call, CUDA events)."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import ips_oracle as O
from ips_b200 import IPSNet, Struct

dev = torch.device('cuda:0')
conf = O.preset('camelyon')
net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
net.load_state_dict(O.make_state(conf, 9, q_gain=8.0))
net.eval()
for B, N in ((1, 50000), (1, 200000), (2, 50000), (16, 50000)):
    x = torch.randn(B, N, 2048, device=dev)
    for shuffle in (True, 'device', False):
        net.shuffle = bool(shuffle)
        net.scan_order_rng = 'device' if shuffle == 'device' else 'reference'
        res = {}
        for mode in ('streamed', 'sequential'):
            net.streamed_select = mode == 'streamed'
            for _ in range(3):
                net.ips(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                net.ips(x)
            e1.record()
            torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / n
        print('B=%d N=%d shuffle=%s  streamed %.3f ms  sequential %.3f ms  status %s' % (
            B, N, shuffle, res['streamed'], res['sequential'], None if net.last_stream_status is None else net.last_stream_status.tolist()), flush=True)
    del x
