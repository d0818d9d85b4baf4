"""Train-step runtime: the grad-mode half of one batch as ONE CUDA graph.

The reference's step (training/iterative.py:124-171) is
    for each loader batch:  mem_patch_iter, mem_pos_iter = net.ips(patches)      # no-grad selection
                            fill_batch(mem_patch, mem_pos, labels, ...)          # copy into the train buffers
    preds = net(mem_patch, mem_pos); loss = compute_loss(...); loss.backward(); optimizer.step()
The second half touches only B x M selected patches: ~500 small kernels whose launch latency, not their work,
sets the step time.  `GraphedTrainStep` captures forward + loss + backward + optimizer step once into a CUDA graph
over static buffers -- `net.ips(..., out=step.buffers, row_offset=n_prep)` writes the winners straight into them
 -- and replays it per batch.  RNG-consuming ops (dropout) stay correct under replay through
PyTorch's graph-safe CUDA generator.
"""
import torch
import torch.nn.functional as F


def compute_loss(conf, preds, labels):
    """Mean over tasks of NLLLoss(log(p + eps)) / BCELoss(p) (training/iterative.py:83-98, criteria main.py:53-61)."""
    loss = 0
    for task in conf.tasks.values():
        p = preds[task['name']].squeeze(-1)
        y = labels[task['name']]
        if task['act_fn'] == 'softmax':
            loss = loss + F.nll_loss(torch.log(p + conf.eps), y)
        else:
            loss = loss + F.binary_cross_entropy(p.view(-1), y.view(-1).to(p.dtype))
    return loss / len(conf.tasks)


def fused_loss(conf, net, mem_patch, mem_pos, labels):
    """forward + heads + loss with the head activation / loss / gradient kernel (`IPSNet.loss`)."""
    return net.loss(mem_patch, mem_pos, labels, conf.eps)


class GraphedTrainStep:
    """forward + loss + backward + optimizer.step of an IPSNet as a replayable CUDA graph.

    buffers   `mem_patch` (B, M, ...), `mem_pos` (B, M, D) or None, `labels` {task: tensor}: static inputs; fill them
              (e.g. `net.ips(x, out=(step.mem_patch, step.mem_pos), row_offset=n_prep)`, `step.labels[k].copy_(...)`)
              and call the object.
    optimizer must be created with `capturable=True` (its step counter lives on the device).
    grad_hook optional callable run after backward (e.g. NCCL all-reduce of the gradients); with a hook the graph covers
              forward + loss + backward and the hook + optimizer step run eagerly after each replay.
    data_parallel  True / a process group: the data-parallel step of north_star.  Every gradient lives in ONE flat fp32
              buffer (`p.grad` are views of it), so the exchange is a single NCCL all-reduce with no flatten / unflatten
              copies, issued between two graphs: [forward + loss + backward] -> all-reduce(flat) -> [1/R scaling +
              optimizer step].  Nothing of the step runs as eager per-tensor kernels.
    """

    def __init__(self, net, conf, optimizer, batch_size, loss_fn=compute_loss, grad_hook=None, warmup=3, fuse_loss=True,
                 data_parallel=None):
        dev = net.device
        self.net, self.conf, self.opt, self.loss_fn, self.grad_hook = net, conf, optimizer, loss_fn, grad_hook
        self.dp_group, self.flat_grad, self.graph_update = None, None, None
        if data_parallel is not None and data_parallel is not False:
            import torch.distributed as dist
            self.dp_group = None if data_parallel is True else data_parallel
            self.dp_world = dist.get_world_size(self.dp_group)
            if grad_hook is not None:
                raise ValueError('give either grad_hook or data_parallel')
            from .distributed import allreduce_gradients
            self.grad_hook = lambda params: allreduce_gradients(params, self.dp_group)      # warm-up steps and the eager fallback
        M = min(net.M, getattr(conf, 'N', net.M)) if getattr(conf, 'N', None) else net.M
        if conf.is_image:
            self.mem_patch = torch.zeros((batch_size, M, conf.n_chan_in, *conf.patch_size), device=dev)
        else:
            self.mem_patch = torch.zeros((batch_size, M, conf.n_chan_in), device=dev)
        self.mem_pos = torch.zeros((batch_size, M, net.D), device=dev) if net.use_pos else None
        self.labels = {}
        for task in conf.tasks.values():
            if task['metric'] == 'multilabel_accuracy':
                self.labels[task['name']] = torch.zeros((batch_size, conf.n_class), dtype=torch.float32, device=dev)
            elif task['act_fn'] == 'sigmoid':
                self.labels[task['name']] = torch.zeros((batch_size,), dtype=torch.float32, device=dev)
            else:
                self.labels[task['name']] = torch.zeros((batch_size,), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), device=dev)
        self.graph = None
        self._warmup = warmup
        self.fuse_loss = fuse_loss and loss_fn is compute_loss      # heads + loss + gradient in one kernel per task

    @property
    def buffers(self):
        return self.mem_patch, self.mem_pos

    def _fwd_bwd(self):
        if self.flat_grad is not None:
            for p in self._with_grad:                        # backward creates the gradients; they are packed below
                p.grad = None
        else:
            # gradients are (re)created by backward: no zero fill and no accumulate launch per parameter; under capture they
            # come from the graph's private pool, at the same addresses in every replay
            self.opt.zero_grad(set_to_none=True)
        if self.fuse_loss:
            loss = fused_loss(self.conf, self.net, self.mem_patch, self.mem_pos, self.labels)
        else:
            loss = self.loss_fn(self.conf, self.net(self.mem_patch, self.mem_pos), self.labels)
        loss.backward()
        if self.flat_grad is not None:
            # pack: a few multi-tensor copies instead of one zero fill + one accumulate launch per parameter; from here on
            # `p.grad` are the views of the flat buffer (what the all-reduce and the optimizer see)
            torch._foreach_copy_(self._flat_views, [p.grad for p in self._with_grad])
            for p, v in zip(self._with_grad, self._flat_views):
                p.grad = v
        self.loss.copy_(loss.detach())

    def _sync_and_update(self):
        if self.grad_hook is not None:
            self.grad_hook([p for p in self.net.parameters() if p.grad is not None])
        self.opt.step()

    def _step(self):
        self._fwd_bwd()
        self._sync_and_update()

    def capture(self, restore_state=True):
        """Warm up on a side stream (lazy workspaces, kernel attributes, autograd buffers, optimizer state), then
        capture.  The warm-up steps are real steps on whatever the buffers hold; with `restore_state` the parameters,
        BatchNorm statistics and optimizer state are put back in place afterwards (the graph keeps their addresses)."""
        dev = self.net.device
        from . import autograd
        if autograd._dist_group(getattr(self.net, 'sync_bn', None)) is not None:
            # synchronised BatchNorm puts collectives inside forward/backward.  With equal shares per rank (fixed per-rank
            # batch: `net.sync_bn_equal_shares = True`) they need no host read and are captured with the rest of the
            # step (NCCL supports stream capture); otherwise the step runs eagerly.
            if not getattr(self.net, 'sync_bn_equal_shares', False):
                self.graph = False
                return self
            autograd.SYNC_BN_EQUAL_SHARES = True
            if autograd.SYNC_BN_PEER is None and getattr(self.net, 'sync_bn_peer', True):
                import torch.distributed as dist
                grp = autograd._dist_group(self.net.sync_bn)
                if dist.get_world_size(grp) <= 8:
                    try:                                     # statistics over NVLink peer memory instead of ~40 NCCL calls per step
                        from .distributed import PeerStatExchange
                        autograd.SYNC_BN_PEER = PeerStatExchange(dev, group=None if grp is dist.group.WORLD else grp)
                    except RuntimeError as e:
                        import warnings
                        warnings.warn('ips_b200: peer-memory exchange unavailable for synchronised BatchNorm (%s): NCCL' % (e,))
        snap = None
        if restore_state:
            snap = ([p.detach().clone() for p in self.net.parameters()], [b.detach().clone() for b in self.net.buffers()],
                    {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for p, st in self.opt.state.items()})
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # with a gradient exchange (NCCL) the graph ends after backward: the collective and the optimizer step run
        # eagerly between replays (a collective inside a capture must be captured identically on every rank)
        if self.dp_group is not None or getattr(self, 'dp_world', None):
            # one flat buffer behind every gradient the warm-up produced (parameters without a gradient stay untouched, as
            # with the reference's optimizer, which skips `grad is None`)
            with_grad = [p for p in self.net.parameters() if p.grad is not None]
            self.flat_grad = torch.zeros(sum(p.numel() for p in with_grad), dtype=torch.float32, device=dev)
            off = 0
            self._with_grad, self._flat_views = with_grad, []
            for p in with_grad:
                n = p.numel()
                self._flat_views.append(self.flat_grad[off:off + n].view_as(p))
                p.grad = self._flat_views[-1]
                off += n
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._fwd_bwd()
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_update):
                self.flat_grad.mul_(1.0 / self.dp_world)
                self.opt.step()
        else:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                if self.grad_hook is None:
                    self._step()
                else:
                    self._fwd_bwd()
        self.net.invalidate_plan()           # warm-up steps and the restore below move the parameters
        if snap is not None:
            with torch.no_grad():
                for p, v in zip(self.net.parameters(), snap[0]):
                    p.copy_(v)
                for b, v in zip(self.net.buffers(), snap[1]):
                    b.copy_(v)
                for p, st in self.opt.state.items():
                    old = snap[2].get(id(p))
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.copy_(old[k]) if old is not None else v.zero_()
        return self

    def __call__(self):
        if self.graph is None:
            self.capture()
        if self.graph is False:
            self._step()
            self.net.invalidate_plan()
            return self.loss
        self.graph.replay()
        if self.graph_update is not None:            # data parallel: ONE all-reduce of the flat gradient between the two graphs
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.dp_group)
            self.graph_update.replay()
        elif self.grad_hook is not None:
            self._sync_and_update()
        self.net.invalidate_plan()           # the replay changed parameters / BatchNorm statistics behind the version counters
        return self.loss
