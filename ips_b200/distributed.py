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
        return ops.topm_stable(ops.scores_from_logits(zc.contiguous(), ca.H, ca.n_token), M)[1]


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
    pad_z = torch.zeros((B, m_max, HT), dtype=zc.dtype, device=dev)
    pad_i = torch.full((B, m_max), -1, dtype=torch.int64, device=dev)
    pad_z[:, :m] = zc
    pad_i[:, :m] = cand_local + lo
    all_z = [torch.empty_like(pad_z) for _ in range(R)]
    all_i = [torch.empty_like(pad_i) for _ in range(R)]
    dist.all_gather(all_z, pad_z, group=group)
    dist.all_gather(all_i, pad_i, group=group)
    counts = [min(M, b - a) if (b - a) > M else (b - a) for a, b in shard_bounds(N, R)]
    zc_all = torch.cat([all_z[r][:, :counts[r]] for r in range(R)], dim=1)
    cand = torch.cat([all_i[r][:, :counts[r]] for r in range(R)], dim=1)      # (B, <= R*M) original indices, rank order
    if cand.shape[1] <= M:
        return cand
    pos = be.merge(zc_all, M)                                                 # identical on every rank
    return torch.gather(cand, 1, pos)


def ips_sharded(net, local_patches, N, group=None, backend=None, mode='exact'):
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
    be = backend or _CudaBackend(net)
    be.pos_offset = lo
    dev = local_patches.device
    if mode == 'merge':
        mem_src = _ips_sharded_merge(net, local_patches, N, group, be, lo, hi)
        return _assemble(net, be, local_patches, mem_src, lo, hi, group)
    if mode != 'exact':
        raise ValueError("mode must be 'exact' or 'merge'")

    # scan order: drawn once (same RNG calls as the reference) on rank 0, broadcast to all ranks
    from .utils import scan_order
    perm, per_inst = None, False
    if net.shuffle:
        if rank == 0:
            perm, per_inst = scan_order(True, net.shuffle_style, B, N, torch.device('cpu'))
        meta = torch.tensor([0 if perm is None else perm.shape[0], int(per_inst)], dtype=torch.int64, device=dev)
        dist.broadcast(meta, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        rows, per_inst = int(meta[0]), bool(meta[1])
        if rows:
            perm = perm.to(dev) if rank == 0 else torch.empty((rows, N), dtype=torch.int64, device=dev)
            dist.broadcast(perm, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        else:
            perm = None

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
