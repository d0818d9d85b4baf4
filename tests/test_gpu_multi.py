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


def _worker(rank, world, port, pre, B, N, over, ret, transport=None, graphed=False):
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
        local = x[:, lo:hi].contiguous()
        if graphed:                                       # the whole sharded call replayed as ONE CUDA graph
            from ips_b200.distributed import ShardedIPS
            sh = ShardedIPS(net, B, N, x.shape[2:], mode='exact').capture(local)
            torch.manual_seed(5)
            mp_s, pos_s = sh(local)
            torch.manual_seed(5)
            mp_s, pos_s = sh(local)                       # a second replay gives the same answer
            ret['status%d' % rank] = sh.ex.status()
        else:
            mp_s, pos_s = ips_sharded(net, local, N, transport=transport)
        mp_s = mp_s.clone()
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
@pytest.mark.parametrize('transport,graphed', [('peer', False), ('nccl', False), ('peer', True)])
def test_sharded_equals_single_gpu(pre, B, N, over, transport, graphed):
    """'exact' schedule == IPSNet.ips on one GPU, bit for bit: NVLink peer-memory kernels (eager and as one CUDA graph)
    and the collective baseline transport."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, pre, B, N, over, ret, transport, graphed), nprocs=2, join=True)
    assert ret[0] and ret[1]
    if graphed:
        assert ret['status0'] == 0 and ret['status1'] == 0


def _worker_merge(rank, world, port, pre, B, N, over, ret, transport=None, output='replicated'):
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
        mem_patch, mem_pos = ips_sharded(net, x[:, lo:hi].contiguous().to(dev), N, mode='merge', transport=transport, output=output)
        ret['idx%d' % rank] = net.last_mem_idx.cpu()
        ret['mem_patch%d' % rank] = mem_patch.cpu()
        if rank == 0:
            ret['mem_patch'] = mem_patch.cpu()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('pre,B,N,over', [('camelyon', 1, 2001, dict(M=200, I=300)),
                                          ('mnist', 2, 100, dict(N=100, M=16, I=20))])
@pytest.mark.parametrize('transport,output', [('peer', 'replicated'), ('nccl', 'replicated'), ('peer', 'batch_split')])
def test_sharded_merge_matches_sharded_oracle(pre, B, N, over, transport, output):
    """P5: local top-M per GPU + exchange of the candidates + one global re-score == the sharded-schedule oracle
    (same block-wise scan order) in fp32; peer-memory kernels (winners replicated or delivered per slide) and the
    collective baseline."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ips_oracle as O
    from ips_b200.distributed import shard_bounds
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_merge, args=(2, port, pre, B, N, over, ret, transport, output), nprocs=2, join=True)
    conf = O.preset(pre, **over)
    sd = O.make_state(conf, 3, q_gain=12.0)
    x = O.make_patches(conf, B, N, 4)
    blocks = []
    for r, (lo, hi) in enumerate(shard_bounds(N, 2)):
        torch.manual_seed(100 + r)
        blocks.append(O.draw_permutation(conf, B, hi - lo) + lo)
    o_patch, _, o_src = O.ips_sharded(sd, conf, x, R=2, perm=torch.cat(blocks, dim=1), tie='stable')
    if output == 'batch_split':                          # slide b lives only on rank b // ceil(B / 2)
        spr = -(-B // 2)
        got_idx = torch.cat([ret['idx%d' % r] for r in range(2)])[:B]
        got_patch = torch.cat([ret['mem_patch%d' % r] for r in range(2)])[:B]
        assert ret['idx0'].shape[0] == min(spr, B)
        assert torch.equal(got_idx, o_src) and torch.equal(got_patch, o_patch)
        return
    assert torch.equal(ret['idx0'], ret['idx1'])
    assert torch.equal(ret['idx0'], o_src)
    assert torch.equal(ret['mem_patch'], o_patch)
    assert torch.equal(ret['mem_patch1'], o_patch)


def _worker_streamed(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import ips_oracle as O
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import ShardedIPS, shard_bounds
        B, N = 1, 24001
        conf = O.preset('camelyon', precision='bf16', M=2000, I=3000)
        sd = O.make_state(conf, 3, q_gain=12.0)
        x = O.make_patches(conf, B, N, 4).to(dev)
        net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
        net.load_state_dict(sd)
        lo, hi = shard_bounds(N, world)[rank]
        local = x[:, lo:hi].contiguous()
        ok = {}
        for rng in ('reference', 'device'):
            net.scan_order_rng = rng
            for mode in ('merge', 'exact'):
                res = {}
                for streamed in (True, False):
                    net.streamed_select = streamed
                    sh = ShardedIPS(net, B, N, x.shape[2:], mode=mode)
                    if sh._gen is not None:
                        sh._gen.manual_seed(11)
                    torch.manual_seed(100 + rank if mode == 'merge' else 11)
                    out, _ = sh(local)
                    res[streamed] = (out.clone(), net.last_mem_idx.clone())
                    if streamed:                              # the same call as ONE graph: replay == eager on the same seed
                        sh.capture(local)
                        if sh._gen is not None:
                            sh._gen.manual_seed(11)
                        torch.manual_seed(100 + rank if mode == 'merge' else 11)
                        out_g, _ = sh(sh.static_input)
                        ok['%s/%s/graph' % (rng, mode)] = (torch.equal(out_g, res[True][0]) and torch.equal(net.last_mem_idx, res[True][1])
                                                          and sh.ex.status() == 0)
                    del sh
                ok['%s/%s' % (rng, mode)] = torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
                if mode == 'exact':                           # == one GPU on the same order
                    torch.manual_seed(11)
                    net.streamed_select = True
                    ref, _ = net.ips(x)
                    ok['%s/exact/single' % rng] = torch.equal(ref, res[True][0]) and torch.equal(net.last_mem_idx, res[True][1])
        ret[rank] = {k: bool(v) for k, v in ok.items()}
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_sharded_streamed_selection_and_device_scan_order():
    """ShardedIPS on bf16 feature bags: the streamed local selection (projector and loop side by side, scan-ordered table)
    gives the winners of the sequential path; scan order drawn with the reference's CPU calls or on the device (inside the
    captured graph; 'exact': one seeded generator per rank, no exchange); graph replay == eager on the same seed; 'exact'
    == IPSNet.ips on one GPU with the same seed."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_streamed, args=(2, port, ret), nprocs=2, join=True)
    for r in range(2):
        bad = [k for k, v in ret[r].items() if not v]
        assert not bad, (r, bad)
        assert len(ret[r]) == 10


def _worker_syncbn(rank, world, port, name, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import ips_oracle as O
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import allreduce_gradients
        from ips_b200.train import compute_loss
        from golden_util import load_case
        z, meta, conf, sd, patches = load_case(name)
        conf = conf.replace(attn_dropout=0.0, dropout=0.0, precision='bf16', sync_bn=True)
        torch.manual_seed(meta['rng_seed'])
        mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
        labels = O.make_labels(conf, meta['B'], meta['label_seed'])
        B = meta['B']
        lo, hi = rank * B // world, (rank + 1) * B // world        # uneven split when B is odd: the statistics weigh by rows
        net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
        net.load_state_dict(sd)
        net.train()
        preds = net(mem_patch[lo:hi].to(dev), None if mem_pos is None else mem_pos[lo:hi].to(dev))
        loss = compute_loss(conf, preds, {k: v[lo:hi].to(dev) for k, v in labels.items()})
        loss.backward()
        allreduce_gradients(list(net.parameters()))
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        if rank == 0:
            ret['loss'] = float(lt) / world
            ret['grads'] = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
            ret['bn'] = {k: v.cpu() for k, v in net.state_dict().items() if 'running' in k}
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('name', ['camelyon_shortcut'])      # even batch: equal shares per rank; fp32-accurate path
def test_syncbn_data_parallel_equals_single_process(name):
    """Data-parallel train step with synchronised BatchNorm: batch split over 2 GPUs + gradient all-reduce ==
    the whole batch on one GPU (loss, every gradient, running statistics), both on the library's kernels."""
    for p in (os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        sys.path.insert(0, p)
    import ips_oracle as O
    from ips_b200 import IPSNet, Struct
    from ips_b200.train import compute_loss
    from golden_util import load_case
    z, meta, conf, sd, patches = load_case(name)
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_syncbn, args=(2, port, name, ret), nprocs=2, join=True)
    conf = conf.replace(attn_dropout=0.0, dropout=0.0, precision='bf16')
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
    labels = O.make_labels(conf, meta['B'], meta['label_seed'])
    dev = torch.device('cuda', 0)
    net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
    net.load_state_dict(sd)
    net.train()
    loss = compute_loss(conf, net(mem_patch.to(dev), None if mem_pos is None else mem_pos.to(dev)), {k: v.to(dev) for k, v in labels.items()})
    loss.backward()
    assert abs(ret['loss'] - loss.item()) <= 2e-2 * max(1.0, abs(loss.item()))
    # both runs compute in bf16: a different fp32 summation order flips bf16 roundings / ReLU gates upstream, so the
    # comparison is per tensor in the L2 sense (the unsynchronised statistics would be off by tens of percent)
    scale = max(float(p.grad.norm()) for p in net.parameters() if p.grad is not None)
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        a, b = p.grad.cpu().double(), ret['grads'][k].double()
        assert float((a - b).norm()) <= 0.15 * float(a.norm()) + 1e-3 * scale, (k, float((a - b).norm()), float(a.norm()))
    for k, v in net.state_dict().items():
        if 'running' in k:
            a, b = v.cpu(), ret['bn'][k]
            assert float((a - b).abs().max()) <= 2e-2 * float(a.abs().max()) + 1e-3, (k, float((a - b).abs().max()))


def _worker_bn_unit(rank, world, port, ret, rows=1001, peer=False):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from ips_b200 import autograd
        from ips_b200.autograd import BatchNormTrainFn
        if peer:                                                # equal shares: statistics over NVLink peer memory, no NCCL call
            from ips_b200.distributed import PeerStatExchange
            autograd.SYNC_BN_EQUAL_SHARES = True
            autograd.SYNC_BN_PEER = PeerStatExchange(dev)
        g = torch.Generator().manual_seed(5)
        cols = 64                                               # rows = 1001: uneven shares (500 / 501)
        x = torch.randn(rows, cols, generator=g) * 2 + 0.5
        dy = torch.randn(rows, cols, generator=g)
        gamma, beta = torch.rand(cols, generator=g) + 0.5, torch.randn(cols, generator=g)
        lo, hi = rank * rows // world, (rank + 1) * rows // world
        xs = x[lo:hi].to(dev).requires_grad_(True)
        gm, bt = gamma.to(dev).requires_grad_(True), beta.to(dev).requires_grad_(True)
        rm, rv = torch.zeros(cols, device=dev), torch.ones(cols, device=dev)
        for _ in range(3 if peer else 1):                       # repeated exchanges alternate the two slots
            if xs.grad is not None:
                xs.grad = None; gm.grad = None; bt.grad = None
                rm.zero_(); rv.fill_(1.0)
            y = BatchNormTrainFn.apply(xs, gm, bt, rm, rv, 0.1, 1e-5, True, True)
            y.backward(dy[lo:hi].to(dev))
        if peer:
            ret['status%d' % rank] = autograd.SYNC_BN_PEER.ex.status()
        dgm, dbt = gm.grad.clone(), bt.grad.clone()
        dist.all_reduce(dgm); dist.all_reduce(dbt)              # parameter gradients are summed by the gradient exchange
        ret[rank] = dict(y=y.detach().cpu(), dx=xs.grad.cpu(), dgamma=dgm.cpu(), dbeta=dbt.cpu(), rm=rm.cpu(), rv=rv.cpu(), lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('rows,peer', [(1001, False), (1000, True)])
def test_synchronised_batchnorm_fn_equals_full_batch(rows, peer):
    """BatchNormTrainFn with a process group on two shares == torch BatchNorm on the whole batch: output, input gradient,
    parameter gradients (summed over ranks), running statistics.  Uneven shares through NCCL collectives; equal shares
    through the peer-memory statistic exchange (the form captured in the data-parallel train step's CUDA graph)."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_bn_unit, args=(2, port, ret, rows, peer), nprocs=2, join=True)
    if peer:
        assert ret['status0'] == 0 and ret['status1'] == 0
    g = torch.Generator().manual_seed(5)
    cols = 64
    x = (torch.randn(rows, cols, generator=g) * 2 + 0.5).requires_grad_(True)
    dy = torch.randn(rows, cols, generator=g)
    gamma = (torch.rand(cols, generator=g) + 0.5).requires_grad_(True)
    beta = torch.randn(cols, generator=g).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(cols, eps=1e-5, momentum=0.1)
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
    y = torch.relu(bn(x))
    y.backward(dy)
    for r in (0, 1):
        lo, hi = ret[r]['lo'], ret[r]['hi']
        torch.testing.assert_close(ret[r]['y'], y[lo:hi].detach(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ret[r]['dx'], x.grad[lo:hi], rtol=1e-3, atol=1e-5)
        torch.testing.assert_close(ret[r]['rm'], bn.running_mean, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(ret[r]['rv'], bn.running_var, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(ret[0]['dgamma'], bn.weight.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(ret[0]['dbeta'], bn.bias.grad, rtol=1e-3, atol=1e-4)


def _worker_dp_step(rank, world, port, name, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import ips_oracle as O
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import allreduce_gradients
        from ips_b200.train import GraphedTrainStep
        from golden_util import load_case
        z, meta, conf, sd, patches = load_case(name)
        conf = conf.replace(attn_dropout=0.0, dropout=0.0, precision='bf16')
        B = meta['B']
        torch.manual_seed(meta['rng_seed'])
        mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
        labels = O.make_labels(conf, B, meta['label_seed'])
        mem_patch = (mem_patch + 0.01 * rank).to(dev)                     # different data on each rank
        out = {}
        for mode in ('hook', 'flat'):
            net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
            net.load_state_dict(sd)
            net.train()
            opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=conf.wd, capturable=True)
            kw = dict(grad_hook=allreduce_gradients) if mode == 'hook' else dict(data_parallel=True)
            step = GraphedTrainStep(net, conf, opt, B, **kw)
            step.mem_patch.copy_(mem_patch)
            if mem_pos is not None:
                step.mem_pos.copy_(mem_pos.to(dev))
            for k, v in labels.items():
                step.labels[k].copy_(v.to(dev))
            step.capture()
            losses = [float(step()) for _ in range(3)]
            out[mode] = (losses, {k: v.detach().float().cpu() for k, v in net.state_dict().items()})
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('name', ['traffic_small', 'camelyon_batch'])
def test_data_parallel_step_flat_gradient_buffer_equals_bucketed_hook(name):
    """The data-parallel train step with ONE all-reduce of a flat gradient buffer between two CUDA graphs takes the same
    optimizer steps as the bucketed all-reduce hook + eager AdamW of round 1, and leaves both ranks with equal weights."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_dp_step, args=(2, port, name, ret), nprocs=2, join=True)
    for r in (0, 1):
        (l_hook, sd_hook), (l_flat, sd_flat) = ret[r]['hook'], ret[r]['flat']
        assert all(abs(a - b) <= 2e-3 * max(1.0, abs(a)) for a, b in zip(l_hook, l_flat)), (l_hook, l_flat)
        for k, v in sd_hook.items():
            if 'running' in k or 'num_batches' in k:
                continue
            assert float((v - sd_flat[k]).abs().max()) <= 2e-3 * (1.0 + float(v.abs().max())), k
    for k, v in ret[0]['flat'][1].items():
        if 'running' in k or 'num_batches' in k:
            continue                                              # per-rank BatchNorm statistics (sync_bn off)
        assert torch.equal(v, ret[1]['flat'][1][k]), k
