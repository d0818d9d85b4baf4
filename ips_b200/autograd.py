"""Autograd glue of the grad-mode train step: `torch.autograd.Function`s whose forward AND backward are the
library's kernels (include/ips_b200.h).  They replace the cuBLAS / ATen calls behind the modules of
`IPSNet.forward` (architecture/ips_net.py:264-283, architecture/transformer.py:85-132) one operator at a
time; an operator without an entry here still runs on PyTorch autograd.

`Linear` -- y = x W^T + b
    forward        NT GEMM   (tcgen05, bf16 operands / fp32 accumulate; `precision='fp32'`: CUDA-core fp32)
    grad input     NN GEMM   dx = dy W
    grad weight    TN GEMM   dW = dy^T x   (MN-major operands, split-K, deterministic)
    grad bias      column sum

`ConvFn` / `StemConvFn` -- bias-free conv2d of the patch encoder on channels-last activations (A12)
    forward        tcgen05 implicit GEMM (the kernels of the selection pass)
    grad input     the same kernel on dy (zero-dilated for stride 2) with flipped, transposed weights
    grad weight    im2col rows (`ipsb_im2col_bf16`) x dy as a TN GEMM (split over the pixels, deterministic)
"""
import os

import torch
from torch import nn

from . import ops


def _pad_cols(t, mult):
    """zero-pad the last dimension of a 2-D tensor up to a multiple of `mult` (tiny class heads)."""
    n = t.shape[1]
    m = (n + mult - 1) // mult * mult
    if m == n:
        return t
    out = torch.zeros((t.shape[0], m), dtype=t.dtype, device=t.device)
    out[:, :n] = t
    return out


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, precision):
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.reshape(-1, K).contiguous().float()
        w = weight.contiguous().float()
        tensor_core = precision == 'bf16' and N % 64 == 0 and K % 8 == 0
        ctx.tensor_core, ctx.in_shape, ctx.has_bias = tensor_core, x.shape, bias is not None
        b = None if bias is None else bias.contiguous().float()
        if tensor_core:
            xb, wb = ops.cast_bf16(x2), ops.cast_bf16(w)
            y = ops.gemm_bf16('nt', xb, wb, shift=b)
            ctx.save_for_backward(xb, wb)
        else:
            if K % 4:                                   # CUDA-core kernel reads float4 along K
                x2, w = _pad_cols(x2, 4), _pad_cols(w, 4)
            y = ops.gemm_f32('nt', x2, w, shift=b)
            ctx.save_for_backward(x2, w)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        xs, ws = ctx.saved_tensors
        N = ws.shape[0]
        dy2 = dy.reshape(-1, N).contiguous().float()
        K = ctx.in_shape[-1]
        dx = dw = db = None
        if ctx.tensor_core:
            dyb = ops.cast_bf16(dy2)
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_bf16('nn', dyb, ws)                              # (M,N) x (N,K)
            if ctx.needs_input_grad[1]:
                if N % 128 == 0:
                    dw = ops.gemm_bf16('tn', dyb, xs)                          # (M,N)^T x (M,K), split over M
                else:
                    dw = ops.gemm_f32('tn', dy2, xs.float())
        else:
            Kp = xs.shape[1]
            if ctx.needs_input_grad[0]:
                dyp, wp = dy2, ws
                if N % 4:                                                        # contraction length must be a multiple of 4
                    dyp = _pad_cols(dy2, 4)
                    wp = torch.zeros((dyp.shape[1], Kp), dtype=ws.dtype, device=ws.device)
                    wp[:N] = ws
                dx = ops.gemm_f32('nn', dyp, wp)[:, :K]
            if ctx.needs_input_grad[1]:
                dw = ops.gemm_f32('tn', dy2 if N % 4 == 0 else _pad_cols(dy2, 4), xs)[:N, :K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dy2)
        if dx is not None:
            dx = dx.reshape(ctx.in_shape)
        return dx, dw, db, None


class Linear(nn.Linear):
    """nn.Linear whose CUDA forward/backward run on the library (same parameters, same state_dict keys)."""
    precision = 'bf16'

    def forward(self, x):
        if not x.is_cuda:
            return super().forward(x)
        return LinearFn.apply(x, self.weight, self.bias, self.precision)


# ---------------------------------------------------------------------------------------------------
# non-GEMM operators of the grad-mode step
# ---------------------------------------------------------------------------------------------------

def _p(t):
    return 0 if t is None else t.data_ptr()


class BatchNormTrainFn(torch.autograd.Function):
    """BatchNorm over rows in batch-statistics mode with an optional fused ReLU (nn.BatchNorm1d + nn.ReLU under
    net.train(), architecture/ips_net.py:58-59; nn.BatchNorm2d on channels-last pixels).  Updates the running
    statistics like nn.BatchNorm.  `group` (a torch.distributed process group, or True for the default group)
    synchronises the statistics over data-parallel ranks: means / variances are combined from every rank's
    (mean, variance, rows) in forward, [sum g, sum g*xhat] are all-reduced in backward, so a batch split over R ranks
    normalises exactly like the whole batch in one process."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu, group=None):
        x = x.contiguous().float()
        rows, cols = x.shape
        mine = torch.empty(2 * cols, dtype=torch.float32, device=x.device)       # [mean | var] of this rank's rows
        mean, var = mine[:cols], mine[cols:]
        scratch = torch.empty(512 * cols, dtype=torch.float32, device=x.device)
        rows_total = rows
        sync = _dist_group(group)
        fused_tail = (running_mean is not None and running_var is not None and running_mean.is_cuda and running_mean.dtype == torch.float32 and running_mean.is_contiguous()
                      and running_var.dtype == torch.float32 and running_var.is_contiguous())
        rstd = None
        if sync is None and fused_tail:                                    # statistics, rstd, running statistics: two launches
            rstd = torch.empty(cols, dtype=torch.float32, device=x.device)
            ops._call('ipsb_bn_stats_finalize_f32', _p(x), _p(mean), _p(var), _p(rstd), _p(running_mean), _p(running_var), float(momentum),
                      rows / max(rows - 1, 1), float(eps), _p(scratch), rows, cols, ops._stream())
        else:
            ops._call('ipsb_bn_stats_finalize_f32', _p(x), _p(mean), _p(var), None, None, None, 0.0, 1.0, float(eps), _p(scratch),
                      rows, cols, ops._stream())
        if sync is not None and SYNC_BN_EQUAL_SHARES:
            # every rank holds the same number of rows (fixed per-rank batch): no row counts to exchange, no host read --
            # the collectives can be captured in the train step's CUDA graph
            import torch.distributed as dist
            R = dist.get_world_size(sync)
            rows_total = rows * R
            if SYNC_BN_PEER is not None and cols % 2 == 0 and fused_tail:  # NVLink peer-memory exchange (no NCCL call):
                # push, then ONE kernel waits, combines, writes mean / rstd and updates the running statistics
                mean, rstd = SYNC_BN_PEER.bn_forward(mine, cols, momentum, rows_total / max(rows_total - 1, 1), eps,
                                                     running_mean, running_var)
            else:
                st = torch.empty((R, 2 * cols), dtype=torch.float32, device=x.device)
                dist.all_gather_into_tensor(st, mine, group=sync)
                mean = st[:, :cols].mean(0)
                var = (st[:, cols:] + (st[:, :cols] - mean) ** 2).mean(0)
        elif sync is not None:
            import torch.distributed as dist
            R = dist.get_world_size(sync)
            mine = torch.cat([mean, var, torch.full((1,), float(rows), device=x.device)])
            parts = [torch.empty_like(mine) for _ in range(R)]
            dist.all_gather(parts, mine, group=sync)
            st = torch.stack(parts)                                        # (R, 2*cols + 1)
            n = st[:, -1:]
            rows_total = int(round(float(n.sum())))
            mean = (st[:, :cols] * n).sum(0) / n.sum()
            var = ((st[:, cols:2 * cols] + (st[:, :cols] - mean) ** 2) * n).sum(0) / n.sum()
        if rstd is None and fused_tail:                                    # rstd + running statistics: one launch
            mean, var = mean.contiguous(), var.contiguous()
            rstd = torch.empty(cols, dtype=torch.float32, device=x.device)
            with torch.no_grad():
                ops._call('ipsb_bn_finalize_f32', _p(mean), _p(var), cols, float(momentum), rows_total / max(rows_total - 1, 1),
                          float(eps), _p(rstd), _p(running_mean), _p(running_var), ops._stream())
        elif rstd is None:
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(mean, alpha=momentum)
                running_var.mul_(1 - momentum).add_(var, alpha=momentum * rows_total / max(rows_total - 1, 1))
            rstd = torch.rsqrt(var + eps)
        y = torch.empty_like(x)
        g, b = gamma.contiguous().float(), beta.contiguous().float()
        ops._call('ipsb_bn_apply_f32', _p(x), _p(mean), _p(rstd), _p(g), _p(b), _p(y), rows, cols, int(relu), ops._stream())
        ctx.save_for_backward(x, y, mean.contiguous(), rstd, g)
        ctx.relu, ctx.sync, ctx.rows_total = relu, sync, rows_total
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, g = ctx.saved_tensors
        rows, cols = x.shape
        dy = dy.contiguous().float()
        sums = torch.empty(2 * cols, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        scratch = torch.empty(512 * cols, dtype=torch.float32, device=x.device)
        ops._call('ipsb_bn_backward_sums_f32', _p(dy), _p(x), _p(y), _p(mean), _p(rstd), _p(sums), _p(scratch), rows, cols,
                  int(ctx.relu), ops._stream())
        all_sums = sums
        if ctx.sync is not None:                                           # dgamma / dbeta stay local (averaged with the other gradients)
            import torch.distributed as dist
            if SYNC_BN_PEER is not None and SYNC_BN_EQUAL_SHARES and cols % 2 == 0:
                all_sums = SYNC_BN_PEER.all_sum(sums)                      # rank order on every rank: identical results
            else:
                all_sums = sums.clone()
                dist.all_reduce(all_sums, op=dist.ReduceOp.SUM, group=ctx.sync)
        ops._call('ipsb_bn_backward_apply_f32', _p(dy), _p(x), _p(y), _p(mean), _p(rstd), _p(g), _p(all_sums), _p(dx), rows, cols,
                  ctx.rows_total, int(ctx.relu), ops._stream())
        return dx, sums[cols:], sums[:cols], None, None, None, None, None, None


# Set by the data-parallel train step when every rank feeds the same number of rows to each BatchNorm (see above).
SYNC_BN_EQUAL_SHARES = False
# `distributed.PeerStatExchange` of the data-parallel group, or None (NCCL collectives)
SYNC_BN_PEER = None


def _dist_group(group):
    """None (no synchronisation) or the process group to synchronise BatchNorm statistics over."""
    if group is None or group is False:
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(None if group is True else group) == 1:
        return None
    return dist.group.WORLD if group is True else group


class LayerNormFn(torch.autograd.Function):
    """LayerNorm over the last dimension, optional affine (nn.LayerNorm, architecture/transformer.py:107,130 and the
    affine-free projector prologue, ips_net.py:56)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous().float()
        D = x2.shape[1]
        if gamma is None:
            y = ops.layernorm_rows(x2, eps)
        else:
            y = ops.residual_layernorm(x2, None, gamma.contiguous().float(), beta.contiguous().float(), eps)
        ctx.save_for_backward(x2, None if gamma is None else gamma.contiguous().float())
        ctx.eps, ctx.shape, ctx.affine = eps, shape, gamma is not None
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, g = ctx.saved_tensors
        rows, D = x2.shape
        dy2 = dy.reshape(rows, D).contiguous().float()
        dx = torch.empty_like(x2)
        xhat = torch.empty_like(x2) if ctx.affine else None
        ops._call('ipsb_layernorm_backward_f32', _p(dy2), _p(x2), _p(g), _p(dx), _p(xhat), rows, D, ctx.eps, ops._stream())
        dg = db = None
        if ctx.affine:
            dg, db = ops.colsum(dy2, xhat), ops.colsum(dy2)
        return dx.view(ctx.shape), dg, db, None


class CrossAttentionFn(torch.autograd.Function):
    """softmax(q k^T) (dropout) v per head and query token (architecture/transformer.py:29-41).  q is already divided by
    the temperature; `mask` (B,H,T,M) of 0/1 floats or None is the attention-dropout keep mask."""

    @staticmethod
    def forward(ctx, q, k, v, mask, keep_scale, H, Dk, Dv):
        q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        B, M = k.shape[:2]
        T = q.shape[0]
        prob = torch.empty((B, H, T, M), dtype=torch.float32, device=k.device)
        out = torch.empty((B, T, H * Dv), dtype=torch.float32, device=k.device)
        chunks = (M + 255) // 256                                   # ipsb_attention_chunks(M)
        scratch = torch.empty(B * H * T * chunks * Dv, dtype=torch.float32, device=k.device)
        ops._call('ipsb_attention_train_fwd_f32', _p(q), _p(k), _p(v), _p(mask), float(keep_scale), _p(prob), _p(out), _p(scratch),
                  B, M, H, Dk, Dv, T, ops._stream())
        ctx.save_for_backward(q, k, v, prob, mask, out)
        ctx.dims = (B, M, H, Dk, Dv, T, float(keep_scale))
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, prob, mask, out = ctx.saved_tensors
        B, M, H, Dk, Dv, T, keep_scale = ctx.dims
        dout = dout.contiguous().float()
        chunks = (M + 255) // 256
        dq_part = torch.empty((B * chunks, T * H * Dk), dtype=torch.float32, device=k.device)
        dk, dv = torch.empty_like(k), torch.empty_like(v)
        ops._call('ipsb_attention_train_bwd_f32', _p(q), _p(k), _p(v), _p(mask), keep_scale, _p(prob), _p(out), _p(dout), _p(dq_part),
                  _p(dk), _p(dv), B, M, H, Dk, Dv, T, ops._stream())
        dq = ops.colsum(dq_part).view(T, H * Dk)
        return dq, dk, dv, None, None, None, None, None


# ----------------------------------------------------------------------------------------
# conv encoder, grad mode (architecture/ips_net.py:17-52 under net.train(): `self.encoder(...)` at :273-274)
# ----------------------------------------------------------------------------------------

_CONST = {}


def _ones_zeros(n, device):
    key = (n, str(device))
    if key not in _CONST:
        _CONST[key] = (torch.ones(n, dtype=torch.float32, device=device), torch.zeros(n, dtype=torch.float32, device=device))
    return _CONST[key]


def _round_up(n, m):
    return (n + m - 1) // m * m


def _conv_wgrad(xb, dyb, kh, kw, stride, pad, Cin=None):
    """dW[co, ci, r, s] = sum_m dy[m, co] * im2col(x)[m, (r, s, ci)] -> (Cout, Cin, kh, kw) fp32 (Cin <= C of xb)."""
    P, H, W, C = xb.shape
    Cout = dyb.shape[-1]
    Cin = C if Cin is None else Cin
    K = kh * kw * C
    dy2 = dyb.reshape(-1, Cout)
    # the long side (K = kh*kw*C) goes to the GEMM's M (multiple of 128), Cout to its N: dW^T = im2col(x)^T dy
    xcol = ops.im2col_bf16(xb, kh, kw, stride, pad, _round_up(K, 128))
    dw_kc = ops.gemm_bf16('tn', xcol, dy2)                               # (Kp, Cout)
    out = torch.empty((Cout, Cin, kh, kw), dtype=torch.float32, device=xb.device)
    ops._call('ipsb_wgrad_to_oihw', _p(dw_kc), Cout, Cin, C, kh, kw, _p(out), ops._stream())
    return out


class ConvFn(torch.autograd.Function):
    """y = conv2d(x, weight, stride, pad) without bias; x, y channels-last (P, H, W, C) fp32, bf16 tensor-core math."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        xb = ops.cast_bf16(x.contiguous().float())
        Cout, Cin, kh, kw = weight.shape
        wf = weight.detach().float().contiguous()
        w_nk = torch.empty((Cout, kh * kw * Cin), dtype=torch.bfloat16, device=x.device)
        w_t = torch.empty((Cin, kh * kw * Cout), dtype=torch.bfloat16, device=x.device) if ctx.needs_input_grad[0] else None
        ops._call('ipsb_conv_weight_layouts', _p(wf), Cout, Cin, kh, kw, _p(w_nk), _p(w_t), ops._stream())   # both operand layouts
        one, zero = _ones_zeros(Cout, x.device)
        y = ops.conv_bf16(xb, w_nk, one, zero, None, Cout, kh, kw, stride, pad, False, 0)
        ctx.save_for_backward(xb, weight, w_t)
        ctx.stride, ctx.pad = stride, pad
        return y.float()

    @staticmethod
    def backward(ctx, dy):
        xb, weight, wt = ctx.saved_tensors
        stride, pad = ctx.stride, ctx.pad
        Cout, Cin, kh, kw = weight.shape
        P, H, W, _ = xb.shape
        Ho, Wo = dy.shape[1], dy.shape[2]
        dyb = ops.cast_bf16(dy.contiguous().float())
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dx = conv(dy zero-dilated to the input grid, weights flipped and transposed, pad k-1-p)
            g = dyb
            if stride != 1:
                g = torch.zeros((P, H, W, Cout), dtype=torch.bfloat16, device=dy.device)
                g[:, ::stride, ::stride][:, :Ho, :Wo] = dyb
            one, zero = _ones_zeros(Cin, dy.device)
            dx = ops.conv_bf16(g, wt, one, zero, None, Cin, kh, kw, 1, kh - 1 - pad, False, 0).float()
        if ctx.needs_input_grad[1]:
            dw = _conv_wgrad(xb, dyb, kh, kw, stride, pad)
        return dx, dw, None, None


class StemConvFn(torch.autograd.Function):
    """The 7x7/2 stem on (P, C, H, W) fp32 patches -> (P, H/2, W/2, 64) fp32 channels-last (no input gradient)."""

    @staticmethod
    def forward(ctx, patches, weight):
        P, C, H, W = patches.shape
        Cout, Cin, kh, kw = weight.shape
        xb = ops.stage_patches(patches.contiguous().float(), P, C, H, W, ops.BF16)          # (P,H,W,4) bf16
        wp = torch.zeros((Cout, 8, 8, 4), dtype=torch.float32, device=patches.device)       # k = r*32 + (s+1)*4 + c
        wp[:, :kh, 1:kw + 1, :Cin] = weight.detach().float().permute(0, 2, 3, 1)
        one, zero = _ones_zeros(Cout, patches.device)
        y = ops.conv_bf16(xb, wp.reshape(Cout, 256).to(torch.bfloat16), one, zero, None, Cout, kh, kw, 2, 3, False, 1)
        ctx.save_for_backward(xb, weight)
        return y.float()

    @staticmethod
    def backward(ctx, dy):
        xb, weight = ctx.saved_tensors
        Cout, Cin, kh, kw = weight.shape
        dw = None
        if ctx.needs_input_grad[1]:
            dyb = ops.cast_bf16(dy.contiguous().float())
            dw = _conv_wgrad(xb, dyb, kh, kw, 2, 3, Cin)
        return None, dw


def _bn2d_train(x, bn, relu, group=None):
    """nn.BatchNorm2d in batch-statistics mode on channels-last activations (rows = pixels)."""
    C = x.shape[-1]
    y = BatchNormTrainFn.apply(x.reshape(-1, C), bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu, group)
    with torch.no_grad():
        bn.num_batches_tracked += 1
    return y.view(x.shape)


class MaxPoolFn(torch.autograd.Function):
    """max_pool2d(3, 2, 1) on channels-last (P, H, W, C) fp32 maps (nn.MaxPool2d of the ResNet stem, ips_net.py:36)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        y = ops.maxpool3x3s2(x, ops.F32)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, = ctx.saved_tensors
        P, H, W, C = x.shape
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        ops._call('ipsb_maxpool3x3s2_bwd_f32', _p(x), _p(dy), _p(dx), P, H, W, C, ops._stream())
        return dx


class AddReluFn(torch.autograd.Function):
    """relu(a + b): the tail of a BasicBlock, one launch forward, one backward (both inputs get the masked gradient)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous().float(), b.contiguous().float()
        y = torch.empty_like(a)
        ops._call('ipsb_add_relu_f32', _p(a), _p(b), _p(y), a.numel(), ops._stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        ops._call('ipsb_relu_bwd_f32', _p(y), _p(dy), _p(dx), y.numel(), ops._stream())
        return dx, dx


def conv_encoder_train(encoder, patches, bn_group=None):
    """Grad-mode forward of the truncated ResNet-18 (`encoder` = the nn.Sequential of ips_net.py:34-50) on
    (P, C, H, W) patches; every convolution and BatchNorm runs forward and backward on the library's kernels.
    Max-pool (forward and backward), residual add + ReLU and the average pool run on the library's kernels too."""
    mods = list(encoder.children())
    x = StemConvFn.apply(patches, mods[0].weight)
    x = _bn2d_train(x, mods[1], True, bn_group)
    if x.shape[-1] % 4 == 0 and os.environ.get('IPS_B200_TORCH_MAXPOOL') is None:
        x = MaxPoolFn.apply(x)
    else:
        x = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()
    for layer in mods[4:-1]:
        for blk in layer:
            idt = x
            y = _bn2d_train(ConvFn.apply(x, blk.conv1.weight, blk.stride, 1), blk.bn1, True, bn_group)
            y = _bn2d_train(ConvFn.apply(y, blk.conv2.weight, 1, 1), blk.bn2, False, bn_group)
            if blk.downsample is not None:
                idt = _bn2d_train(ConvFn.apply(x, blk.downsample[0].weight, blk.stride, 0), blk.downsample[1], False, bn_group)
            x = AddReluFn.apply(y, idt) if y.numel() % 4 == 0 else torch.relu(y + idt)
    return x.mean(dim=(1, 2))


# ----------------------------------------------------------------------------------------
# heads + loss (get_preds, ips_net.py:157-166, + compute_loss, training/iterative.py:83-98) in one kernel
# ----------------------------------------------------------------------------------------

class HeadLossFn(torch.autograd.Function):
    """loss = NLLLoss(log(softmax(z) + eps), y) or BCELoss(sigmoid(z), t), mean reduction; the kernel also writes
    d loss / d z, so backward is a scale."""

    @staticmethod
    def forward(ctx, z, labels, act_fn, eps):
        z2 = z.reshape(z.shape[0], -1).contiguous().float()
        rows, n = z2.shape
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z2)
        if act_fn == 'softmax':
            cls = labels.reshape(rows).contiguous().to(torch.int64)
            ops._call('ipsb_head_loss_f32', _p(z2), _p(cls), 0, rows, n, 0, float(eps), _p(loss), _p(dz), 0, ops._stream())
        else:
            tgt = labels.reshape(rows, n).contiguous().float()
            ops._call('ipsb_head_loss_f32', _p(z2), 0, _p(tgt), rows, n, 1, float(eps), _p(loss), _p(dz), 0, ops._stream())
        ctx.save_for_backward(dz)
        ctx.shape = z.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return (dz * g).view(ctx.shape), None, None, None
