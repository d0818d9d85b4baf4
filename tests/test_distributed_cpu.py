"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharded selection assembles the same
result as the single-process oracle, ragged all-gather, gradient averaging.  The CUDA kernels are
replaced by oracle-based stand-ins through the `backend` hook; the collectives are real."""
import math
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ips_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


class OracleBackend:
    """CPU stand-in for the CUDA kernels, built from the oracle's functions."""

    def __init__(self, sd, conf):
        self.sd, self.conf, self.pos_offset = sd, conf, 0

    def logits(self, local):
        B, n = local.shape[:2]
        emb = O.encode(self.sd, self.conf, local.reshape(B * n, *local.shape[2:])).view(B, n, -1)
        if self.conf.use_pos:
            emb = emb + O.pos_table(self.conf.D, self.conf.N)[self.pos_offset:self.pos_offset + n]
        return O.attn_logits(self.sd, self.conf, emb).permute(0, 3, 1, 2).reshape(B, n, -1).contiguous()

    def select(self, z, perm, per_instance):
        conf = self.conf
        B, N, _ = z.shape
        out = []
        for b in range(B):
            order = torch.arange(N) if perm is None else perm[b if per_instance else 0]
            mem = torch.arange(conf.M)
            for it in range(math.ceil((N - conf.M) / conf.I)):
                lo = conf.M + it * conf.I
                cand = torch.cat([mem, torch.arange(lo, min(lo + conf.I, N))])
                zz = z[b, order[cand]].view(1, -1, conf.H, conf.n_token).permute(0, 2, 3, 1)
                sc = torch.softmax(zz, -1).mean(1).transpose(1, 2).mean(-1)[0]
                final = cand[torch.sort(sc, descending=True, stable=True)[1][:conf.M]]
                mem = final.sort()[0]
            out.append(order[final])
        return torch.stack(out)

    def merge(self, zc, M):
        conf = self.conf
        zz = zc.view(zc.shape[0], -1, conf.H, conf.n_token).permute(0, 2, 3, 1)
        sc = torch.softmax(zz, -1).mean(1).transpose(1, 2).mean(-1)
        return torch.sort(sc, dim=-1, descending=True, stable=True)[1][:, :M]

    def gather(self, local, local_idx):
        rows = [torch.stack([local[b, i] if i >= 0 else torch.zeros_like(local[b, 0]) for i in local_idx[b].tolist()])
                for b in range(local.shape[0])]
        return torch.stack(rows)


def _worker_merge(rank, world, port, case, ret, M_over=None):
    """north_star schedule: local top-M per rank, all-gather of the candidates, one global re-score."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import ips_sharded, shard_bounds
        from golden_util import load_case
        z, meta, conf, sd, patches = load_case(case)
        if M_over:
            conf = conf.replace(M=M_over)
        N = patches.shape[1]
        net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
        net.load_state_dict(sd)
        lo, hi = shard_bounds(N, world)[rank]
        torch.manual_seed(100 + rank)                           # every rank shuffles its own slice
        mem_patch, mem_pos = ips_sharded(net, patches[:, lo:hi].contiguous(), N, backend=OracleBackend(sd, conf), mode='merge')
        ret['idx%d' % rank] = net.last_mem_idx
        if rank == 0:
            ret.update(mem_patch=mem_patch, mem_pos=mem_pos)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case,M_over', [('camelyon_batch', None), ('mnist_small', None),
                                         ('mnist_small', 20)])       # slices (18 patches) not longer than M: keep everything
def test_sharded_merge_schedule_world2(case, M_over):
    """P5: the local-top-M + merge schedule equals the sharded-schedule oracle run in one process."""
    from golden_util import load_case
    from ips_b200.distributed import shard_bounds
    z, meta, conf, sd, patches = load_case(case)
    if M_over:
        conf = conf.replace(M=M_over)
    B, N = patches.shape[:2]
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_merge, args=(2, port, case, ret, M_over), nprocs=2, join=True)
    blocks = []
    for r, (lo, hi) in enumerate(shard_bounds(N, 2)):           # the equivalent block-wise permutation
        torch.manual_seed(100 + r)
        blocks.append(O.draw_permutation(conf, B, hi - lo) + lo)
    perm = torch.cat(blocks, dim=1)
    o_patch, o_pos, o_src = O.ips_sharded(sd, conf, patches, R=2, perm=perm, tie='stable')
    assert torch.equal(ret['idx0'], ret['idx1'])
    assert torch.equal(ret['idx0'], o_src)
    assert torch.equal(ret['mem_patch'], o_patch)
    if conf.use_pos:
        assert torch.equal(ret['mem_pos'], o_pos)


def _worker(rank, world, port, case, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import ips_sharded, shard_bounds, gather_logit_table, allreduce_gradients
        from golden_util import load_case
        z, meta, conf, sd, patches = load_case(case)
        N = patches.shape[1]
        net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
        net.load_state_dict(sd)
        lo, hi = shard_bounds(N, world)[rank]
        torch.manual_seed(77)                                   # only rank 0's draw is used
        mem_patch, mem_pos = ips_sharded(net, patches[:, lo:hi].contiguous(), N, backend=OracleBackend(sd, conf))
        # ragged all-gather
        t = torch.full((2, hi - lo, 3), float(rank))
        full = gather_logit_table(t, N)
        ok_gather = all(bool((full[:, a:b] == r).all()) for r, (a, b) in enumerate(shard_bounds(N, world)))
        # gradient averaging
        p = torch.nn.Parameter(torch.zeros(5))
        p.grad = torch.full((5,), float(rank + 1))
        allreduce_gradients([p])
        ok_grad = bool(torch.allclose(p.grad, torch.full((5,), (world + 1) / 2)))
        if rank == 0:
            ret.update(mem_patch=mem_patch, mem_pos=mem_pos, idx=net.last_mem_idx, ok_gather=ok_gather, ok_grad=ok_grad)
        else:
            ret['idx1'] = net.last_mem_idx
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case', ['camelyon_batch', 'mnist_small'])
def test_sharded_selection_world2(case):
    from golden_util import load_case
    z, meta, conf, sd, patches = load_case(case)
    N = patches.shape[1]
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, case, ret), nprocs=2, join=True)
    torch.manual_seed(77)
    perm = O.draw_permutation(conf, patches.shape[0], N)
    o_patch, o_pos, o_src = O.ips(sd, conf, patches, perm=perm, tie='stable')
    assert torch.equal(ret['idx'], ret['idx1'])                  # every rank selects the same patches
    assert torch.equal(ret['idx'], o_src)                        # ... and they are the single-process selection
    assert torch.equal(ret['mem_patch'], o_patch)                # all-reduce assembly is exact
    if conf.use_pos:
        assert torch.equal(ret['mem_pos'], o_pos)
    assert ret['ok_gather'] and ret['ok_grad']
