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


def ips_sharded(net, local_patches, N, group=None, backend=None):
    """Sequence-sharded `IPSNet.ips`.  `local_patches` is this rank's slice (B, n_r, ...) of the patch axis in
    rank order (`shard_bounds(N, world)`).  Returns (mem_patch, mem_pos) on every rank, identical to
    `net.ips` on the full tensor with the same scan order."""
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
