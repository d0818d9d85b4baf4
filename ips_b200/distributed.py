"""Multi-GPU execution of the IPS hot path (one process per GPU, torch.distributed / NCCL).

Two modes:

* images of a batch are independent -> split the batch over ranks, no collective
  (``bench.py --gpus N`` runs this; nothing to do here);
* ONE long patch sequence sharded over ranks (`ips_sharded`): every rank encodes its
  contiguous slice into logits, the slices are all-gathered into the full (B, N, H*T) logit
  table (N*H*T*4 bytes: 1.6 MB for a 50k-patch slide), and every rank runs the same
  selection loop on it.  Because a patch's logits do not depend on the other patches, the
  result is IDENTICAL to the single-GPU `IPSNet.ips` on the concatenated sequence -- unlike a
  local-top-M + candidate merge, which changes the softmax context.  The selected
  patches are then assembled with one all-reduce (each rank contributes the rows it owns).

  `mode='merge'` is the scalable schedule of `north_star` / SURVEY 8e: every rank runs the selection loop on ITS slice
  only (N/R patches, local top-M), the R*M candidates' logits and indices are all-gathered (M*(H*T*4+8) bytes per rank)
  and every rank re-scores them in one buffer and keeps the global stable top-M.  It matches the sharded-schedule
  oracle (`oracle/ips_oracle.py::ips_sharded`, parity P5) exactly, and the sequential scan whenever H*n_token == 1;
  with several heads it differs from the sequential scan in a fraction of a percent of the picks (F5).

Transports.  On CUDA the exchange runs through NVLink PEER MEMORY (`ShardedIPS`, `PeerExchange`; csrc/peer.cu): every
rank owns an exchange buffer whose CUDA-IPC handle is all-gathered ONCE at set-up; per call the candidates (logits ‖ index,
one packed record per candidate), the logit slices and the winning patches are written straight into the peers' buffers by
one kernel each, ordered by epoch flags in the buffers' headers -- no NCCL call on the data path, no torch glue, no host
round trip, so the whole call can be captured in a CUDA graph (`ShardedIPS.capture`).  A rank copies only the winner rows
it OWNS (1/R of M rows) to the ranks that need them; with `output='batch_split'` slide b goes only to the rank whose
data-parallel train step consumes it.  The collective transport (`transport='nccl'`: ONE packed all-gather of the
candidates, winners by all-reduce) is kept as the baseline, for CPU tensors (gloo tests) and for slices not longer than M.

`allreduce_gradients` is the data-parallel gradient exchange of the train step.
"""
import torch
import torch.distributed as dist


def shard_bounds(N, R):
    """Contiguous slices of the patch axis; sizes differ by at most one."""
    base, rem = divmod(N, R)
    out, lo = [], 0
    for r in range(R):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class _CudaBackend:
    """The product kernels (ips_b200.ops); tests on CPU inject an oracle-based stand-in."""

    def __init__(self, net):
        self.net = net

    def logits(self, local_patches):
        return self.net.patch_logits(local_patches, pos_offset=self.pos_offset)

    def select(self, z, perm, per_instance):
        from . import ops
        ca = self.net.transf.crs_attn
        return ops.select_loop(z.contiguous(), perm, per_instance, ca.H, ca.n_token, self.net.M, self.net.I)[1]

    def gather(self, local_patches, local_idx):
        from . import ops
        return ops.gather_rows(local_patches.contiguous(), local_idx.contiguous(), local_patches.shape[1])

    def merge(self, zc, M):
        """(B, L, HT) candidate logits -> positions (B, M) of the stable top-M of their scores in ONE buffer."""
        from . import ops
        ca = self.net.transf.crs_attn
        return ops.merge_candidates(zc.contiguous(), ca.H, ca.n_token, M)


def gather_logit_table(z_local, N, group=None):
    """All-gather per-rank (B, n_r, HT) slices (n_r may differ by one) into (B, N, HT)."""
    R = dist.get_world_size(group)
    B, n_r, HT = z_local.shape
    n_max = -(-N // R)
    buf = z_local
    if n_r < n_max:                                            # pad to a common length
        buf = torch.zeros((B, n_max, HT), dtype=z_local.dtype, device=z_local.device)
        buf[:, :n_r] = z_local
    parts = [torch.empty_like(buf) for _ in range(R)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    out = torch.empty((B, N, HT), dtype=z_local.dtype, device=z_local.device)
    for r, (lo, hi) in enumerate(shard_bounds(N, R)):
        out[:, lo:hi] = parts[r][:, :hi - lo]
    return out


def local_scan_order(net, B, n_local, device):
    """This rank's scan order of its own slice in 'merge' mode: the reference's shuffle (utils/utils.py:33-58) applied
    to the slice.  The equivalent single-process permutation is cat_r(lo_r + local_order_r) (block-wise shuffle)."""
    from .utils import scan_order
    return scan_order(net.shuffle, net.shuffle_style, B, n_local, device)


def _ips_sharded_merge(net, local_patches, N, group, be, lo, hi):
    R = dist.get_world_size(group)
    B, n = local_patches.shape[:2]
    M, dev = net.M, local_patches.device
    perm, per_inst = local_scan_order(net, B, n, torch.device('cpu'))
    if perm is not None:
        perm = perm.to(dev).contiguous()
    z_local = be.logits(local_patches)                                        # (B, n, HT), true positions via pos_offset
    if n > M:
        cand_local = be.select(z_local, perm, per_inst)                       # (B, M) local indices, best first
    elif perm is None:
        cand_local = torch.arange(n, device=dev).unsqueeze(0).expand(B, -1).contiguous()
    else:                                                                     # short slice: everything, in scan order
        cand_local = (perm if per_inst else perm[:1].expand(B, -1)).contiguous()
    m = cand_local.shape[1]
    HT = z_local.shape[-1]
    zc = be.gather(z_local, cand_local)                                       # (B, m, HT): the candidates' logits
    # all-gather the candidates (slices not longer than M contribute fewer than M): pad to M rows
    m_max = min(M, -(-N // R))
    # ONE collective: a packed record per candidate = HT fp32 logits ‖ the int64 index as two 32-bit words
    packed = torch.zeros((B, m_max, HT + 2), dtype=torch.float32, device=dev)
    packed[:, :m, :HT] = zc
    packed[:, :m, HT:] = (cand_local + lo).unsqueeze(-1).contiguous().view(torch.float32)
    gathered = torch.empty((R, B, m_max, HT + 2), dtype=torch.float32, device=dev)
    if dev.type == 'cuda':
        dist.all_gather_into_tensor(gathered, packed, group=group)
    else:
        dist.all_gather(list(gathered.unbind(0)), packed, group=group)
    counts = [min(M, b - a) if (b - a) > M else (b - a) for a, b in shard_bounds(N, R)]
    zc_all = torch.cat([gathered[r, :, :counts[r], :HT] for r in range(R)], dim=1)
    cand = torch.cat([gathered[r, :, :counts[r], HT:].contiguous().view(torch.int64).squeeze(-1) for r in range(R)], dim=1)
    # (B, <= R*M) original indices, rank order
    if cand.shape[1] <= M:
        return cand
    pos = be.merge(zc_all, M)                                                 # identical on every rank
    return torch.gather(cand, 1, pos)


def _align(v, a=256):
    return (v + a - 1) // a * a


class PeerExchange:
    """One exchange buffer per rank of `group` (same node), every peer's buffer mapped into this process through CUDA
    IPC (csrc/peer.cu).  The set-up is collective: one all-gather of the 64-byte handles, one barrier."""

    def __init__(self, nbytes, device, group=None):
        from . import ops
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError('ips_b200: the peer-memory exchange spans one node (at most 8 ranks)')
        self.device = torch.device(device)
        self.nbytes = _align(int(nbytes), 4096)
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)       # header = zeros: epoch 0 everywhere
        torch.cuda.synchronize(self.device)
        handle, offset = ops.peer_export(self.buf)
        mine = torch.tensor(list(handle) + list(int(offset).to_bytes(8, 'little')), dtype=torch.uint8, device=self.device)
        every = torch.empty((self.world, mine.numel()), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine, group=group)
        every = every.cpu()
        bases = []
        for q in range(self.world):
            if q == self.rank:
                bases.append(self.buf.data_ptr())
                continue
            rec = bytes(every[q].tolist())
            bases.append(ops.peer_open(rec[:64], int.from_bytes(rec[64:72], 'little'), self.device.index))
        self.ctx = ops.peer_ctx(self.rank, self.world, bases)
        dist.barrier(group=group)                     # nobody signals into a buffer that is not zeroed and mapped yet

    def view(self, offset, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.buf[offset:offset + nbytes].view(dtype).view(shape)

    def status(self):
        from . import ops
        return ops.peer_status(self.ctx)


class PeerStatExchange:
    """All-gathers of a few KB over NVLink peer memory: the statistic exchanges of synchronised BatchNorm in the
    data-parallel train step (forward: every rank's mean / variance; backward: [sum g, sum g*xhat]).  Two kernels per
    exchange (push + flag, wait + copy) instead of an NCCL collective; capturable in the step's CUDA graph."""

    def __init__(self, device, max_floats=4096, group=None):
        self.R = dist.get_world_size(group)
        self.stride = _align(self.R * max_floats * 4, 16)
        self.max_floats = max_floats
        self.ex = PeerExchange(4096 + 2 * self.stride, device, group)

    def all_gather(self, v):
        """v: contiguous fp32 vector (length a multiple of 4, at most max_floats) -> (R, len) in rank order."""
        from . import ops
        if v.numel() % 4 or v.numel() > self.max_floats:
            raise ValueError('PeerStatExchange: vector of %d floats' % v.numel())
        return ops.peer_allgather_small(self.ex.ctx, v.contiguous(), 4096, self.stride, 2, self.R)

    def _check(self, v):
        if v.numel() % 4 or v.numel() > self.max_floats or not v.is_contiguous():
            raise ValueError('PeerStatExchange: contiguous vector of %d floats' % v.numel())

    def bn_forward(self, mean_var, cols, momentum, unbias, eps, running_mean, running_var):
        """[mean | var] of this rank -> (mean, rstd) over all ranks (equal rows per rank); running statistics updated."""
        from . import ops
        self._check(mean_var)
        return ops.peer_bn_forward(self.ex.ctx, mean_var, cols, 4096, self.stride, 2, momentum, unbias, eps, running_mean, running_var)

    def all_sum(self, v):
        """Sum of `v` over the ranks, in rank order (bit-identical on every rank)."""
        from . import ops
        self._check(v)
        return ops.peer_allgather_sum(self.ex.ctx, v, 4096, self.stride, 2)


class ShardedIPS:
    """Sequence-sharded `IPSNet.ips` over NVLink peer memory for fixed shapes: this rank's slice is
    (B, n_local, *row_shape) fp32 of a sequence of N patches (`shard_bounds(N, world)[rank]`).

        sharded = ShardedIPS(net, B, N, row_shape, mode='merge')      # collective set-up
        mem_patch, mem_pos = sharded(local_patches)                   # every call: kernels only
        sharded.capture(local_patches)                                # optional: the same call as ONE CUDA graph

    mode 'merge' (north_star): local selection loop -> push M candidates (logits ‖ index) to every rank -> wait ->
    global re-score + stable top-M -> push the owned winner rows -> wait.  mode 'exact': push the logit slice to every
    rank -> wait -> replicated loop -> push the owned winner rows -> wait.
    output 'replicated': every rank returns all B slides (like `net.ips` on every rank); 'batch_split': slide b is
    delivered only to rank b // ceil(B / world) (the rank whose data-parallel train step consumes it).
    The returned tensors are views of this rank's exchange buffer: valid until the next call."""

    def __init__(self, net, B, N, row_shape, mode='merge', group=None, output='replicated'):
        if mode not in ('merge', 'exact'):
            raise ValueError("mode must be 'exact' or 'merge'")
        if output not in ('replicated', 'batch_split'):
            raise ValueError("output must be 'replicated' or 'batch_split'")
        self.net, self.mode, self.group, self.output = net, mode, group, output
        self.R, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.B, self.N, self.M = int(B), int(N), int(net.M)
        self.row_shape = tuple(row_shape)
        self.lo, self.hi = shard_bounds(self.N, self.R)[self.rank]
        if min(b - a for a, b in shard_bounds(self.N, self.R)) <= self.M:
            raise ValueError('ShardedIPS needs every slice longer than M (use the collective transport otherwise)')
        ca = net.transf.crs_attn
        self.HT = ca.H * ca.n_token
        row_bytes = 4
        for d in self.row_shape:
            row_bytes *= d
        if row_bytes % 16:
            raise ValueError('ShardedIPS: patch rows must be multiples of 16 bytes')
        self.row_bytes = row_bytes
        self.L = self.R * self.M
        self.spr = -(-self.B // self.R) if output == 'batch_split' else 0           # slides per rank
        B_out = self.spr if self.spr else self.B
        off = 4096
        self.cz_off = self.ci_off = self.z_off = 0
        if mode == 'merge':
            self.cz_off = off; off = _align(off + self.B * self.L * self.HT * 4)
            self.ci_off = off; off = _align(off + self.B * self.L * 8)
        else:
            self.z_off = off; off = _align(off + self.B * self.N * self.HT * 4)
        self.out_off = off; off = _align(off + B_out * self.M * row_bytes)
        self.ex = PeerExchange(off, net.device, group)
        self.mem_src = torch.empty((self.B, self.M), dtype=torch.int64, device=net.device)
        self.out = self.ex.view(self.out_off, (B_out, self.M, *self.row_shape), torch.float32)
        if mode == 'merge':
            self.cz = self.ex.view(self.cz_off, (self.B, self.L, self.HT), torch.float32)
            self.ci = self.ex.view(self.ci_off, (self.B, self.L), torch.int64)
        else:
            self.z = self.ex.view(self.z_off, (self.B, self.N, self.HT), torch.float32)
        self.graph = None
        self.bytes_pushed_per_call = 0          # filled by the first call: payload bytes this rank stores into OTHER ranks
        # conf.scan_order_rng == 'device': the scan order is drawn on the GPU inside the call (part of the captured graph).
        # 'exact' mode needs the SAME order on every rank: a dedicated generator per rank, seeded with one broadcast seed,
        # draws it locally -- no exchange.  'merge' shuffles each slice independently (default CUDA generator).
        self._gen = None
        if mode == 'exact' and net.shuffle and getattr(net, 'scan_order_rng', 'reference') == 'device':
            seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(net.device)
            dist.broadcast(seed, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            self._gen = torch.Generator(device=net.device)
            self._gen.manual_seed(int(seed.item()))

    # ---- scan order (host side, outside any graph) ---------------------------------------------------------
    def _draw_order(self):
        net, dev = self.net, self.net.device
        if net.shuffle and getattr(net, 'scan_order_rng', 'reference') == 'device':
            return 'device', False                    # drawn inside _run
        if self.mode == 'merge':                      # every rank shuffles its own slice (block-wise shuffle, local_scan_order)
            return net._draw_scan_order(self.B, self.hi - self.lo, torch.device('cpu'), dev)
        return _broadcast_scan_order(net, self.B, self.N, dev, self.group)

    # ---- the kernels of one call (capturable) -----------------------------------------------------------------
    def _run(self, local_patches, perm, per_inst):
        from . import ops
        net, ca, ctx = self.net, self.net.transf.crs_attn, self.ex.ctx
        M = self.M
        n = self.hi - self.lo
        plan = net._get_plan()
        if isinstance(perm, str):                     # device-side scan order (see __init__)
            from .utils import scan_order
            perm, per_inst = scan_order(True, net.shuffle_style, self.B, n if self.mode == 'merge' else self.N, None,
                                        'device', net.device, self._gen)
            perm = None if perm is None else perm.contiguous()
        if (self.mode == 'merge' and net.streamed_select and not net.is_image and plan.get('p_table') is not None
                and ops.streamed_select_ok(self.B, n, ca.H * ca.n_token, M, net.I)):
            # projector and local loop side by side (ops.projector_select); the logit table stays in scan order
            pos, cand, _, zs, _ = ops.projector_select(local_patches.reshape(self.B * n, -1), plan['p_w'], plan['p_table'],
                                                       self.B, n, perm, per_inst, ca.H, ca.n_token, M, net.I)
            ops.peer_push_candidates(ctx, zs, cand, self.lo, self.L, self.rank * M, self.cz_off, self.ci_off, rows=pos)
            ops.peer_wait(ctx, 0)
            win = ops.merge_candidates(self.cz, ca.H, ca.n_token, M)
            ci = self.ci
            z_local = None
        else:
            z_local = net.patch_logits(local_patches, pos_offset=self.lo)              # (B, n, HT), true positions
        if z_local is None:
            pass
        elif self.mode == 'merge':
            cand = ops.select_loop(z_local.contiguous(), perm, per_inst, ca.H, ca.n_token, M, net.I)[1]   # local, best first
            ops.peer_push_candidates(ctx, z_local, cand, self.lo, self.L, self.rank * M, self.cz_off, self.ci_off)
            ops.peer_wait(ctx, 0)
            win = ops.merge_candidates(self.cz, ca.H, ca.n_token, M)                          # positions in the candidate list
            ci = self.ci
        else:
            ops.peer_push_logits(ctx, z_local, self.N, self.lo, self.z_off)
            ops.peer_wait(ctx, 0)
            win = ops.select_loop(self.z, perm, per_inst, ca.H, ca.n_token, M, net.I)[1]      # global indices
            ci = None
        ops.peer_push_winners(ctx, local_patches, self.lo, win, ci, self.L, self.row_bytes, self.spr, self.out_off, self.mem_src)
        ops.peer_wait(ctx, 1)
        mem_pos = ops.gather_rows(net.pos_enc[0].contiguous(), self._my_rows(self.mem_src), 0) if net.use_pos else None
        return mem_pos

    def _my_rows(self, t):
        if not self.spr:
            return t
        return t[self.rank * self.spr: min(self.B, (self.rank + 1) * self.spr)].contiguous()

    def _result(self, mem_pos):
        self.net.last_mem_idx = self._my_rows(self.mem_src)
        n_out = self.net.last_mem_idx.shape[0]
        return self.out[:n_out], mem_pos

    def _check(self, local_patches):
        shape = (self.B, self.hi - self.lo, *self.row_shape)
        if tuple(local_patches.shape) != shape or local_patches.dtype != torch.float32 or not local_patches.is_cuda:
            raise ValueError('ShardedIPS: expected a CUDA fp32 slice of shape %s, got %s %s' % (shape, tuple(local_patches.shape), local_patches.dtype))
        if not local_patches.is_contiguous():
            raise ValueError('ShardedIPS: the slice must be contiguous')

    @torch.no_grad()
    def __call__(self, local_patches):
        self._check(local_patches)
        perm, per_inst = self._draw_order()
        if self.graph is None:
            return self._result(self._run(local_patches, perm, per_inst))
        if local_patches.data_ptr() != self._g_in.data_ptr():     # (fill `static_input` in place to avoid this copy)
            self._g_in.copy_(local_patches, non_blocking=True)
        if perm is not None and not isinstance(perm, str):
            self._g_perm.copy_(perm, non_blocking=True)
        self.graph.replay()
        return self._result(self._g_pos)

    @property
    def static_input(self):
        """The graph's input buffer (after `capture`): write the next slice into it and call the object with it."""
        return self._g_in

    @torch.no_grad()
    def capture(self, local_patches):
        """Capture one call (encoder, selection, both exchanges) as ONE CUDA graph over static buffers; later calls copy
        the slice and the scan order into them and replay.  Collective: every rank must capture."""
        self._check(local_patches)
        perm, per_inst = self._draw_order()
        self._g_in = local_patches.clone()
        self._g_perm, self._g_inst = (perm if (perm is None or isinstance(perm, str)) else perm.clone()), per_inst
        self._run(self._g_in, self._g_perm, per_inst)                     # warm-up: plans, workspaces, function attributes
        torch.cuda.synchronize(self.net.device)
        dist.barrier(group=self.group)
        g = torch.cuda.CUDAGraph()
        if self._gen is not None:
            g.register_generator_state(self._gen)
        side = torch.cuda.Stream(device=self.net.device)
        side.wait_stream(torch.cuda.current_stream(self.net.device))
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self._g_pos = self._run(self._g_in, self._g_perm, per_inst)
        torch.cuda.current_stream(self.net.device).wait_stream(side)
        self.graph = g
        return self


def _broadcast_scan_order(net, B, N, dev, group):
    """'exact' mode: the scan order is drawn once (same RNG calls as the reference) on rank 0 and broadcast."""
    from .utils import scan_order
    rank = dist.get_rank(group)
    perm, per_inst = None, False
    if not net.shuffle:
        return None, False
    src = dist.get_global_rank(group, 0) if group is not None else 0
    if rank == 0:
        perm, per_inst = scan_order(True, net.shuffle_style, B, N, torch.device('cpu'))
    meta = torch.tensor([0 if perm is None else perm.shape[0], int(per_inst)], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=src, group=group)
    rows, per_inst = int(meta[0]), bool(meta[1])
    if not rows:
        return None, False
    perm = perm.to(dev) if rank == 0 else torch.empty((rows, N), dtype=torch.int64, device=dev)
    dist.broadcast(perm, src=src, group=group)
    return perm.contiguous(), per_inst


def _peer_path_ok(net, local_patches, N, group, backend):
    if backend is not None or not local_patches.is_cuda or local_patches.dtype != torch.float32:
        return False
    R = dist.get_world_size(group)
    if R > 8 or min(b - a for a, b in shard_bounds(N, R)) <= net.M:
        return False
    row_bytes = 4
    for d in local_patches.shape[2:]:
        row_bytes *= d
    return row_bytes % 16 == 0


def ips_sharded(net, local_patches, N, group=None, backend=None, mode='exact', transport=None, output='replicated'):
    """Sequence-sharded `IPSNet.ips`.  `local_patches` is this rank's slice (B, n_r, ...) of the patch axis in
    rank order (`shard_bounds(N, world)`).  Returns (mem_patch, mem_pos) on every rank.
    mode 'exact': identical to `net.ips` on the full tensor with the same scan order (logit table all-gathered, loop
    replicated).  mode 'merge': local top-M per rank + one global re-score of the R*M candidates (the loop shards)."""
    R = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(N, R)[rank]
    B = local_patches.shape[0]
    assert local_patches.shape[1] == hi - lo, 'slice length does not match shard_bounds'
    M = net.M
    if M >= N:
        raise ValueError('M >= N: nothing to select, use net.ips on the gathered sequence')
    if mode not in ('exact', 'merge'):
        raise ValueError("mode must be 'exact' or 'merge'")
    explicit = transport is not None
    if transport is None:
        transport = 'peer' if _peer_path_ok(net, local_patches, N, group, backend) else 'nccl'
    if transport == 'peer':                                       # NVLink peer-memory kernels, set up once per shape
        key = (mode, output, B, N, tuple(local_patches.shape[2:]), id(group))
        cache = net.__dict__.setdefault('_sharded', {})
        if key not in cache:
            try:
                cache[key] = ShardedIPS(net, B, N, local_patches.shape[2:], mode=mode, group=group, output=output)
            except RuntimeError as e:                             # e.g. an allocator whose blocks cannot be IPC-exported
                if explicit:
                    raise
                import warnings
                warnings.warn('ips_b200: peer-memory exchange unavailable (%s); using the collective transport' % (e,))
                cache[key] = None
        if cache[key] is not None:
            return cache[key](local_patches.contiguous())
        transport = 'nccl'
    if transport != 'nccl':
        raise ValueError("transport must be 'peer' or 'nccl'")
    if output != 'replicated':
        raise ValueError("output='batch_split' needs the peer-memory transport")
    be = backend or _CudaBackend(net)
    be.pos_offset = lo
    dev = local_patches.device
    if mode == 'merge':
        mem_src = _ips_sharded_merge(net, local_patches, N, group, be, lo, hi)
        return _assemble(net, be, local_patches, mem_src, lo, hi, group)
    perm, per_inst = _broadcast_scan_order(net, B, N, dev, group)

    z_local = be.logits(local_patches)                               # (B, n_r, HT)
    z = gather_logit_table(z_local, N, group)                         # (B, N, HT) everywhere
    mem_src = be.select(z, perm, per_inst)                            # (B, M) original indices, identical on all ranks
    return _assemble(net, be, local_patches, mem_src, lo, hi, group)


def _assemble(net, be, local_patches, mem_src, lo, hi, group):
    """Winners (original indices, identical on all ranks) -> (mem_patch, mem_pos): every rank contributes its rows."""
    net.last_mem_idx = mem_src
    owned = (mem_src >= lo) & (mem_src < hi)
    local_idx = torch.where(owned, mem_src - lo, torch.full_like(mem_src, -1))      # -1 -> zero row
    mem_patch = be.gather(local_patches, local_idx)
    dist.all_reduce(mem_patch, op=dist.ReduceOp.SUM, group=group)    # exactly one rank contributes each row
    mem_pos = None
    if net.use_pos:
        mem_pos = net.pos_enc[0][mem_src]
    return mem_patch, mem_pos


def allreduce_gradients(params, group=None, bucket_mb=64):
    """Average the gradients of a data-parallel train step across ranks (NCCL all-reduce on flat buckets)."""
    R = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0
    limit = bucket_mb * (1 << 20)

    def flush():
        if not bucket:
            return
        flat = torch._utils._flatten_dense_tensors(bucket)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(R)
        for g, f in zip(bucket, torch._utils._unflatten_dense_tensors(flat, bucket)):
            g.copy_(f)
        bucket.clear()

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= limit:
            flush()
            size = 0
    flush()
