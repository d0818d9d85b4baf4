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
