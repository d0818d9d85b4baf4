"""2-GPU NCCL test: sequence-sharded selection equals the single-GPU selection bit for bit.
Skipped on boxes with one GPU (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, pre, B, N, over, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import ips_oracle as O
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import ips_sharded, shard_bounds
        conf = O.preset(pre, precision='bf16', **over)
        sd = O.make_state(conf, 3, q_gain=12.0)
        x = O.make_patches(conf, B, N, 4).to(dev)
        net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
        net.load_state_dict(sd)
        lo, hi = shard_bounds(N, world)[rank]
        torch.manual_seed(5)
        mp_s, pos_s = ips_sharded(net, x[:, lo:hi].contiguous(), N)
        idx_s = net.last_mem_idx.clone()
        torch.manual_seed(5)
        mp_1, pos_1 = net.ips(x)
        ok = torch.equal(idx_s, net.last_mem_idx) and torch.equal(mp_s, mp_1)
        if pos_1 is not None:
            ok = ok and torch.equal(pos_s, pos_1)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('pre,B,N,over', [('camelyon', 1, 20001, dict(M=2000, I=3000)),
                                          ('mnist', 2, 100, dict(N=100, M=16, I=20))])
def test_sharded_equals_single_gpu(pre, B, N, over):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, pre, B, N, over, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _worker_merge(rank, world, port, pre, B, N, over, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import ips_oracle as O
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import ips_sharded, shard_bounds
        conf = O.preset(pre, precision='fp32', **over)
        sd = O.make_state(conf, 3, q_gain=12.0)
        x = O.make_patches(conf, B, N, 4)
        net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
        net.load_state_dict(sd)
        lo, hi = shard_bounds(N, world)[rank]
        torch.manual_seed(100 + rank)                       # every rank shuffles its own slice
        mem_patch, mem_pos = ips_sharded(net, x[:, lo:hi].contiguous().to(dev), N, mode='merge')
        ret['idx%d' % rank] = net.last_mem_idx.cpu()
        if rank == 0:
            ret['mem_patch'] = mem_patch.cpu()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('pre,B,N,over', [('camelyon', 1, 2001, dict(M=200, I=300)),
                                          ('mnist', 2, 100, dict(N=100, M=16, I=20))])
def test_sharded_merge_matches_sharded_oracle(pre, B, N, over):
    """P5: local top-M per GPU + all-gather of the candidates + one global re-score == the sharded-schedule oracle
    (same block-wise scan order) in fp32."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ips_oracle as O
    from ips_b200.distributed import shard_bounds
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_merge, args=(2, port, pre, B, N, over, ret), nprocs=2, join=True)
    conf = O.preset(pre, **over)
    sd = O.make_state(conf, 3, q_gain=12.0)
    x = O.make_patches(conf, B, N, 4)
    blocks = []
    for r, (lo, hi) in enumerate(shard_bounds(N, 2)):
        torch.manual_seed(100 + r)
        blocks.append(O.draw_permutation(conf, B, hi - lo) + lo)
    o_patch, _, o_src = O.ips_sharded(sd, conf, x, R=2, perm=torch.cat(blocks, dim=1), tie='stable')
    assert torch.equal(ret['idx0'], ret['idx1'])
    assert torch.equal(ret['idx0'], o_src)
    assert torch.equal(ret['mem_patch'], o_patch)
