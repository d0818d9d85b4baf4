"""Autograd glue of the grad-mode train step: `torch.autograd.Function`s whose forward AND backward are the
library's kernels (include/ips_b200.h).  They replace the cuBLAS / ATen calls behind the modules of
`IPSNet.forward` (architecture/ips_net.py:264-283, architecture/transformer.py:85-132) one operator at a
time; an operator without an entry here still runs on PyTorch autograd.

`Linear` -- y = x W^T + b
    forward        NT GEMM   (tcgen05, bf16 operands / fp32 accumulate; `precision='fp32'`: CUDA-core fp32)
    grad input     NN GEMM   dx = dy W
    grad weight    TN GEMM   dW = dy^T x   (MN-major operands, split-K, deterministic)
    grad bias      column sum
"""
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
    net.train(), architecture/ips_net.py:58-59).  Updates the running statistics like nn.BatchNorm1d."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu):
        x = x.contiguous().float()
        rows, cols = x.shape
        mean = torch.empty(cols, dtype=torch.float32, device=x.device)
        var = torch.empty_like(mean)
        scratch = torch.empty(128 * cols, dtype=torch.float32, device=x.device)
        ops._call('ipsb_bn_stats_f32', _p(x), _p(mean), _p(var), _p(scratch), rows, cols, ops._stream())
        with torch.no_grad():
            running_mean.mul_(1 - momentum).add_(mean, alpha=momentum)
            running_var.mul_(1 - momentum).add_(var, alpha=momentum * rows / max(rows - 1, 1))
        rstd = torch.rsqrt(var + eps)
        y = torch.empty_like(x)
        g, b = gamma.contiguous().float(), beta.contiguous().float()
        ops._call('ipsb_bn_apply_f32', _p(x), _p(mean), _p(rstd), _p(g), _p(b), _p(y), rows, cols, int(relu), ops._stream())
        ctx.save_for_backward(x, y, mean, rstd, g)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, g = ctx.saved_tensors
        rows, cols = x.shape
        dy = dy.contiguous().float()
        sums = torch.empty(2 * cols, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        scratch = torch.empty(128 * cols, dtype=torch.float32, device=x.device)
        ops._call('ipsb_bn_backward_f32', _p(dy), _p(x), _p(y), _p(mean), _p(rstd), _p(g), _p(sums), _p(dx), _p(scratch), rows, cols,
                  int(ctx.relu), ops._stream())
        return dx, sums[cols:], sums[:cols], None, None, None, None, None


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
        ops._call('ipsb_attention_train_fwd_f32', _p(q), _p(k), _p(v), _p(mask), float(keep_scale), _p(prob), _p(out),
                  B, M, H, Dk, Dv, T, ops._stream())
        ctx.save_for_backward(q, k, v, prob, mask)
        ctx.dims = (B, M, H, Dk, Dv, T, float(keep_scale))
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, prob, mask = ctx.saved_tensors
        B, M, H, Dk, Dv, T, keep_scale = ctx.dims
        dout = dout.contiguous().float()
        dq_part = torch.empty((B, T * H * Dk), dtype=torch.float32, device=k.device)
        dk, dv = torch.empty_like(k), torch.empty_like(v)
        ops._call('ipsb_attention_train_bwd_f32', _p(q), _p(k), _p(v), _p(mask), keep_scale, _p(prob), _p(dout), _p(dq_part),
                  _p(dk), _p(dv), B, M, H, Dk, Dv, T, ops._stream())
        dq = ops.colsum(dq_part).view(T, H * Dk)
        return dq, dk, dv, None, None, None, None, None
