"""Tensor-level wrappers over the C ABI (include/ips_b200.h).

Each function validates its tensors, passes raw device pointers and the current
CUDA stream to the library, and returns freshly allocated outputs.  The main
entry points are also registered as PyTorch custom ops under ``torch.ops.ips_b200``
(see the bottom of the file).  No function here has a non-CUDA implementation.
"""
import ctypes
import os

import torch

from . import _lib

F32, BF16 = 0, 1


LAUNCHES = 0          # kernels launched through the C ABI by this process (bench.py reports it)
TIMER = None          # optional {kernel name: [(start_event, end_event, tag)]} filled when set to a dict
TAG = None            # free-form tag attached to timed launches (e.g. layer name + algorithmic FLOPs)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _call(name, *args):
    """One C-ABI call = one kernel launch on the current stream."""
    global LAUNCHES
    fn = getattr(_lib.load(), name)
    LAUNCHES += 1
    if TIMER is None:
        _lib.check(fn(*args))
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(fn(*args))
    e1.record()
    TIMER.setdefault(name, []).append((e0, e1, TAG))


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f'ips_b200: {name} must be a CUDA tensor (no CPU path exists)')
    if t.dtype != dtype:
        raise RuntimeError(f'ips_b200: {name} must be {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise RuntimeError(f'ips_b200: {name} must be contiguous')


def _dt(code):
    return torch.bfloat16 if code == BF16 else torch.float32


# ------------------------------------------------------------------ staging / movement

def stage_patches(src, n_rows, C, H, W, dt, row_idx=None, first_row=0, cpad=4):
    """(rows,C,H,W) fp32 -> (n_rows,H,W,cpad) channels-last of dtype code `dt`."""
    _chk(src, torch.float32, 'src')
    _chk(row_idx, torch.int64, 'row_idx')
    out = torch.empty((n_rows, H, W, cpad), dtype=_dt(dt), device=src.device)
    _call('ipsb_stage_patches', _p(src), _p(row_idx), first_row, n_rows, C, H, W, cpad, dt,
                                              _p(out), _stream())
    return out


def stage_patches_padded(src, n_rows, C, H, W, row_idx=None, first_row=0):
    """(rows,C,H,W) fp32 -> zero-bordered (n_rows,H+6,W+6,4) bf16 frame for the TMA stem (image at row 3, col 4)."""
    _chk(src, torch.float32, 'src')
    _chk(row_idx, torch.int64, 'row_idx')
    out = torch.empty((n_rows, H + 6, W + 6, 4), dtype=torch.bfloat16, device=src.device)
    _call('ipsb_stage_patches_padded', _p(src), _p(row_idx), first_row, n_rows, C, H, W, 3, 4, H + 6, W + 6,
          _p(out), _stream())
    return out


def stage_patches_s2d(src, n_rows, C, H, W, row_idx=None, first_row=0):
    """(rows,C,H,W) fp32 -> space-to-depth frame (n_rows * (H/2+3)*(W/2+3), 16) bf16 for the shifted-window stem."""
    _chk(src, torch.float32, 'src')
    _chk(row_idx, torch.int64, 'row_idx')
    Sp = (H // 2 + 3) * (W // 2 + 3)
    out = torch.empty((n_rows * Sp, 16), dtype=torch.bfloat16, device=src.device)
    _call('ipsb_stage_patches_s2d', _p(src), _p(row_idx), first_row, n_rows, C, H, W, _p(out), _stream())
    return out


def conv_stem_s2d(frame, w_nk, scale, shift, P, H, W, Cout=64, relu=True):
    """7x7/2 stem on the s2d frame; returns the wide-row-order output (P * Sp, Cout) bf16."""
    _chk(frame, torch.bfloat16, 'frame'); _chk(w_nk, torch.bfloat16, 'w')
    Sp = (H // 2 + 3) * (W // 2 + 3)
    y = torch.empty((P * Sp, Cout), dtype=torch.bfloat16, device=frame.device)
    _call('ipsb_conv_bf16_umma', _p(frame), _p(w_nk), _p(scale), _p(shift), 0, _p(y), P, H, W, 16, Cout, 7, 7, 2, 3, int(relu), 4,
          _stream())
    return y


def im2col_bf16(x, kh, kw, stride, pad, Kp=None):
    """(P,H,W,C) bf16 channels-last -> (P*Ho*Wo, Kp) bf16 im2col rows, column (r*kw+s)*C + c, zero padded to Kp."""
    _chk(x, torch.bfloat16, 'x')
    P, H, W, C = x.shape
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    K = kh * kw * C
    Kp = K if Kp is None else Kp
    out = torch.empty((P * Ho * Wo, Kp), dtype=torch.bfloat16, device=x.device)
    _call('ipsb_im2col_bf16', _p(x), _p(out), P, H, W, C, kh, kw, stride, pad, Kp, _stream())
    return out


def stem_pool_s2d(frame, w_nk, scale, shift, P, H, W, relu=True, out=None):
    """Fused 7x7/2 stem + BN + ReLU + 3x3/2 max-pool on the s2d frame -> padded-flat (rows, 64) bf16."""
    _chk(frame, torch.bfloat16, 'frame'); _chk(w_nk, torch.bfloat16, 'w')
    Hq, Wq = (H // 2 - 1) // 2 + 1, (W // 2 - 1) // 2 + 1
    if out is None:
        out = torch.zeros((pf_geo(P, Hq, Wq)[0], 64), dtype=torch.bfloat16, device=frame.device)
    _call('ipsb_stem_pool_s2d', _p(frame), _p(w_nk), _p(scale), _p(shift), _p(out), P, H, W, int(relu), _stream())
    return out


def maxpool3x3s2_pf_strided(x, P, H, W, C, in_Wp, in_Sp, out=None):
    """max-pool of a (P, H, W, C) map stored with row pitch in_Wp / patch pitch in_Sp (in pixels) -> PF (rows, C)."""
    _chk(x, torch.bfloat16, 'x')
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if out is None:
        out = torch.zeros((pf_geo(P, Ho, Wo)[0], C), dtype=torch.bfloat16, device=x.device)
    _call('ipsb_maxpool3x3s2_pf_strided', _p(x), _p(out), P, H, W, C, in_Wp, in_Sp, _stream())
    return out


def gather_rows(src, idx, batch_stride_rows, out=None):
    """dst[b,m] = src_rows[b*batch_stride_rows + idx[b,m]]; src viewed as rows of src.shape[-k:].
    `out`: contiguous (B, M, *row_shape) destination, e.g. a row range of the train-step buffer."""
    _chk(idx, torch.int64, 'idx')
    if not src.is_cuda or not src.is_contiguous():
        raise RuntimeError('ips_b200: gather source must be a contiguous CUDA tensor')
    B, M = idx.shape
    row_shape = src.shape[2:] if batch_stride_rows else src.shape[1:]
    row_bytes = src.element_size()
    for s in row_shape:
        row_bytes *= s
    if out is None:
        out = torch.empty((B, M, *row_shape), dtype=src.dtype, device=src.device)
    elif tuple(out.shape) != (B, M, *row_shape) or out.dtype != src.dtype or not out.is_contiguous() or out.device != src.device:
        raise RuntimeError('ips_b200: gather destination must be a contiguous %s tensor on the source device' % ((B, M, *row_shape),))
    _call('ipsb_gather_rows', _p(src), batch_stride_rows, _p(idx), B, M, row_bytes, _p(out), _stream())
    return out


def maxpool3x3s2(x, dt):
    P, H, W, C = x.shape
    _chk(x, _dt(dt), 'x')
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((P, Ho, Wo, C), dtype=x.dtype, device=x.device)
    _call('ipsb_maxpool3x3s2', _p(x), _p(y), P, H, W, C, dt, _stream())
    return y


def avgpool(x, dt):
    P, H, W, C = x.shape
    _chk(x, _dt(dt), 'x')
    y = torch.empty((P, C), dtype=torch.float32, device=x.device)
    _call('ipsb_avgpool', _p(x), _p(y), P, H * W, C, dt, _stream())
    return y


def layernorm_rows(x, eps):
    _chk(x, torch.float32, 'x')
    y = torch.empty_like(x)
    _call('ipsb_layernorm_rows_f32', _p(x), _p(y), x.shape[0], x.shape[1], eps, _stream())
    return y


def rows_to_bf16(x, layernorm, eps=1e-5):
    """(rows, F) fp32 or bf16 features -> bf16 GEMM operand, optionally LayerNorm'ed (no affine) in fp32."""
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if x.dtype == torch.bfloat16:
        _chk(x, torch.bfloat16, 'x')
        _call('ipsb_rows_bf16_to_bf16', _p(x), _p(y), x.shape[0], x.shape[1], int(layernorm), eps, _stream())
        return y
    _chk(x, torch.float32, 'x')
    _call('ipsb_rows_to_bf16', _p(x), _p(y), x.shape[0], x.shape[1], int(layernorm), eps, _stream())
    return y


# ------------------------------------------------------------------ encoder layers

def conv_f32(x, w_kc, scale, shift, res, Cout, kh, kw, stride, pad, relu):
    """x (P,H,W,Cin) fp32 NHWC, w_kc (kh*kw*Cin, Cout) fp32."""
    _chk(x, torch.float32, 'x'); _chk(w_kc, torch.float32, 'w'); _chk(res, torch.float32, 'res')
    P, H, W, Cin = x.shape
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    y = torch.empty((P, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    _call('ipsb_conv_f32', _p(x), _p(w_kc), _p(scale), _p(shift), _p(res), _p(y), P, H, W, Cin, Cout,
                                         kh, kw, stride, pad, int(relu), _stream())
    return y


def conv_bf16(x, w_nk, scale, shift, res, Cout, kh, kw, stride, pad, relu, mode=0):
    """x (P,H,W,Cin) bf16 NHWC, w_nk (Cout, Kpad) bf16 K-major; tcgen05 implicit GEMM."""
    _chk(x, torch.bfloat16, 'x'); _chk(w_nk, torch.bfloat16, 'w'); _chk(res, torch.bfloat16, 'res')
    P, H, W, Cin = x.shape
    if mode == 3:                                    # zero-bordered frame: the padding is in the data
        Ho, Wo = (H - kh) // stride + 1, (W - kw) // stride + 1
    else:
        Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    y = torch.empty((P, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
    _call('ipsb_conv_bf16_umma', _p(x), _p(w_nk), _p(scale), _p(shift), _p(res), _p(y), P, H, W, Cin,
                                               Cout, kh, kw, stride, pad, int(relu), mode, _stream())
    return y


# ---- padded-flat (PF) activations: see ips_b200/csrc/pf.cuh ------------------------------------

def pf_geo(P, H, W):
    """(rows, G0, Wp, Sp) of the PF layout of P patches of HxW pixels."""
    Wp, Sp, G0 = W + 1, (H + 1) * (W + 1), W + 2
    return (G0 + P * Sp + Wp + 2 + 7) // 8 * 8, G0, Wp, Sp


def to_pf(x):
    """dense (P,H,W,C) -> PF (rows, C) with zero pads (torch ops; test / debugging helper)."""
    P, H, W, C = x.shape
    rows, G0, Wp, Sp = pf_geo(P, H, W)
    out = torch.zeros((rows, C), dtype=x.dtype, device=x.device)
    body = out[G0:G0 + P * Sp].view(P, H + 1, Wp, C)
    body[:, :H, :W] = x
    return out


def from_pf(x, P, H, W):
    """PF (rows, C) -> dense (P,H,W,C) view copy (test / debugging helper)."""
    rows, G0, Wp, Sp = pf_geo(P, H, W)
    return x[G0:G0 + P * Sp].view(P, H + 1, Wp, -1)[:, :H, :W].contiguous()


def conv_bf16_pf(x, w_nk, scale, shift, res, P, H, W, Cout, kh, kw, stride, pad, relu, in_pf, out_pf, out=None):
    """bf16 tcgen05 convolution reading / writing padded-flat tensors.  x: PF (rows,Cin) if in_pf else dense
    (P,H,W,Cin).  Returns PF (rows,Cout) if out_pf else dense (P,Ho,Wo,Cout).  3x3/1 PF->PF with
    Cout in {64,128} runs the shifted-window kernel.  `out` (PF) must have zero pad rows."""
    _chk(x, torch.bfloat16, 'x'); _chk(w_nk, torch.bfloat16, 'w'); _chk(res, torch.bfloat16, 'res')
    Cin = x.shape[-1]
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    if out is None:
        if out_pf:
            out = torch.zeros((pf_geo(P, Ho, Wo)[0], Cout), dtype=torch.bfloat16, device=x.device)
        else:
            out = torch.empty((P, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
    _call('ipsb_conv_bf16_pf', _p(x), _p(w_nk), _p(scale), _p(shift), _p(res), _p(out), P, H, W, Cin, Cout, kh, kw,
          stride, pad, int(relu), int(in_pf), int(out_pf), _stream())
    return out


def maxpool3x3s2_pf(x, out=None):
    """dense bf16 (P,H,W,C) -> PF (rows,C) of the pooled (Ho,Wo) map."""
    _chk(x, torch.bfloat16, 'x')
    P, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if out is None:
        out = torch.zeros((pf_geo(P, Ho, Wo)[0], C), dtype=torch.bfloat16, device=x.device)
    _call('ipsb_maxpool3x3s2_pf', _p(x), _p(out), P, H, W, C, _stream())
    return out


def avgpool_pf(x, P, H, W):
    _chk(x, torch.bfloat16, 'x')
    C = x.shape[-1]
    y = torch.empty((P, C), dtype=torch.float32, device=x.device)
    _call('ipsb_avgpool_pf', _p(x), _p(y), P, H, W, C, _stream())
    return y


def linear_f32(a, w, scale=None, shift=None, relu=False):
    """y = act((a @ w.T) * scale + shift); a (M,K), w (N,K) fp32."""
    _chk(a, torch.float32, 'a'); _chk(w, torch.float32, 'w')
    M, K = a.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _call('ipsb_linear_f32', _p(a), _p(w), _p(scale), _p(shift), _p(y), M, N, K, int(relu), _stream())
    return y


def linear_bf16(a, w, scale=None, shift=None, relu=False):
    """tcgen05 GEMM: a (M,K) bf16, w (N,K) bf16 -> fp32 (M,N)."""
    _chk(a, torch.bfloat16, 'a'); _chk(w, torch.bfloat16, 'w')
    M, K = a.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _call('ipsb_linear_bf16_umma', _p(a), _p(w), _p(scale), _p(shift), _p(y), M, N, K, int(relu), _stream())
    return y


# ---- "bf16x3": fp32 values as hi + lo bf16, products as three bf16 MMAs accumulated in fp32 (csrc/split3.cu) -------------
# Activations in the "tri" layout (pixels, 3C) = [hi | hi | lo]; weights [wh | wl | wh] along the input channels.

def tri_from_f32(x):
    """fp32 (rows..., C) -> tri bf16 (rows..., 3C) = [hi | hi | lo] with hi + lo = x to ~17 significant bits."""
    _chk(x, torch.float32, 'x')
    C = x.shape[-1]
    tri = torch.empty((*x.shape[:-1], 3 * C), dtype=torch.bfloat16, device=x.device)
    _call('ipsb_sum3_split', _p(x), 0, 0, 0, 0, 0, 0, _p(tri), 0, 0, x.numel(), C, _stream())
    return tri


def conv_f32out(x, w, scale, shift, Cout, kh, kw, stride, pad, mode=0):
    """bf16 (P,H,W,Cin) x bf16 (Cout, K) -> fp32 (P,Ho,Wo,Cout) = conv * scale + shift (no activation); fp32 accumulation."""
    _chk(x, torch.bfloat16, 'x'); _chk(w, torch.bfloat16, 'w')
    P, H, W, Cin = x.shape
    if mode == 3:
        Ho, Wo = (H - kh) // stride + 1, (W - kw) // stride + 1
    else:
        Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    y = torch.empty((P, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    _call('ipsb_conv_bf16_f32out', _p(x), _p(w), _p(scale), _p(shift), _p(y), P, H, W, Cin, Cout, kh, kw, stride, pad, mode, _stream())
    return y


def sum_split_tri(parts, relu, C, res_f32=None, res_tri=None, want_f32=False, want_tri=True):
    """a (+ b + c) (+ residual: fp32 or tri) (+ ReLU) -> tri tensor (..., 3C) and / or the fp32 sum."""
    a = parts[0]
    b = parts[1] if len(parts) > 1 else None
    c = parts[2] if len(parts) > 2 else None
    tri = torch.empty((*a.shape[:-1], 3 * C), dtype=torch.bfloat16, device=a.device) if want_tri else None
    out = torch.empty_like(a) if want_f32 else None
    _call('ipsb_sum3_split', _p(a), _p(b), _p(c), _p(res_f32), _p(res_tri), 0, int(relu), _p(tri), 0, _p(out), a.numel(), C, _stream())
    return tri, out


def sum3_maxpool_tri(parts, relu=True):
    """stem tail: a + b + c -> ReLU -> max_pool2d(3, 2, 1) -> tri (P, Ho, Wo, 3C)."""
    a, b, c = parts
    P, H, W, C = a.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    tri = torch.empty((P, Ho, Wo, 3 * C), dtype=torch.bfloat16, device=a.device)
    _call('ipsb_sum3_maxpool_split', _p(a), _p(b), _p(c), P, H, W, C, int(relu), _p(tri), 0, 1, _stream())
    return tri


def stage_patches_padded_split(src, n_rows, C, H, W, first_row=0):
    """(rows,C,H,W) fp32 -> hi / lo zero-bordered (n_rows,H+6,W+6,4) bf16 frames for the TMA stem."""
    _chk(src, torch.float32, 'src')
    hi = torch.empty((n_rows, H + 6, W + 6, 4), dtype=torch.bfloat16, device=src.device)
    lo = torch.empty_like(hi)
    _call('ipsb_stage_patches_padded_split', _p(src), first_row, n_rows, C, H, W, _p(hi), _p(lo), _stream())
    return hi, lo


def fold_plan(items_dev, n_items, blocks_per_item):
    """Refill every folded convolution weight / BatchNorm scale+shift described by the device item table in ONE launch."""
    _chk(items_dev, torch.uint8, 'items')
    _call('ipsb_fold_plan', _p(items_dev), n_items, blocks_per_item, _stream())


def projector_table(scale, shift, w_bf16, U):
    """(N, 12) fp32 column table of ipsb_projector_logits: a_n, b_n = a_n * colsum(w_bf16)_n, shift_n, 0, U[n, 0..7]."""
    N, HT = U.shape
    a = scale.float()
    cs = w_bf16.float().sum(dim=1)
    tab = torch.zeros((N, 12), dtype=torch.float32, device=U.device)
    tab[:, 0], tab[:, 1], tab[:, 2] = a, a * cs, shift.float()
    tab[:, 4:4 + HT] = U
    return tab.contiguous()


def projector_logits(x, w_bf16, table, HT, eps=1e-5):
    """Feature rows (rows, K) fp32 or bf16 -> logits (rows, HT) fp32 in ONE kernel: LayerNorm (algebraic) + Linear on the
    tensor cores + BatchNorm(eval) + ReLU + score projection (ips_net.py:54-60, transformer.py:71-83)."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError('ips_b200: features must be fp32 or bf16')
    _chk(x, x.dtype, 'x'); _chk(w_bf16, torch.bfloat16, 'w'); _chk(table, torch.float32, 'table')
    rows, K = x.shape
    N = w_bf16.shape[0]
    z = torch.empty((rows, HT), dtype=torch.float32, device=x.device)
    _call('ipsb_projector_logits', _p(x), int(x.dtype == torch.bfloat16), _p(w_bf16), _p(table), _p(z), rows, K, N, HT, eps, _stream())
    return z


def make_resnet_desc(plan, dt, D, HT):
    """Pack the folded-parameter plan of IPSNet into the ipsb_resnet_desc the native executor reads."""
    def conv(e):
        return _lib.ConvDesc(_p(e['w']), _p(e['scale']), _p(e['shift']), e['cin'], e['cout'], e['kh'], e['kw'],
                             e['stride'], e['pad'], e['mode'], 0)
    d = _lib.ResnetDesc()
    d.dt, d.n_blocks, d.stem = dt, len(plan['blocks']), conv(plan['stem'])
    for i, b in enumerate(plan['blocks']):
        d.blocks[i].c1, d.blocks[i].c2 = conv(b['c1']), conv(b['c2'])
        d.blocks[i].has_ds = int(b['ds'] is not None)
        if b['ds'] is not None:
            d.blocks[i].ds = conv(b['ds'])
    d.D, d.HT, d.U, d.add_tab = D, HT, _p(plan['U']), _p(plan['posU'])
    return d


def image_geo(images, patch_size, patch_stride):
    """ipsb_image_geo + patches per image for (B, C, Himg, Wimg) images cut like the reference's `unfold` calls."""
    Himg, Wimg = images.shape[-2:]
    ph, pw = patch_size
    sh, sw = patch_stride
    n_rows, n_cols = (Himg - ph) // sh + 1, (Wimg - pw) // sw + 1
    return _lib.ImageGeo(Himg, Wimg, sh, sw, n_rows * n_cols), n_rows * n_cols


def gather_patches_image(images, geo, idx, patch_size, out=None):
    """(B, M, C, ph, pw) fp32 <- patch idx[b, m] of image b; idx None = every patch in grid order (on-device patchify)."""
    _chk(images, torch.float32, 'images'); _chk(idx, torch.int64, 'idx')
    B, C = images.shape[:2]
    M = geo.n_per_image if idx is None else idx.shape[1]
    if out is None:
        out = torch.empty((B, M, C, *patch_size), dtype=torch.float32, device=images.device)
    _call('ipsb_gather_patches_image', _p(images), ctypes.byref(geo), _p(idx), B, M, C, patch_size[0], patch_size[1], _p(out), _stream())
    return out


def resnet_logits(desc, patches, n_per_image, chunk, workspace_cache, want_emb=False, first_row=0, n_rows=None, z=None, lanes=1,
                  geo=None, patch_size=None, src_first_row=0):
    """Whole eval-mode encoder + logit projection of (rows,C,H,W) fp32 patches in ONE library call.
    `first_row` / `n_rows` restrict the call to a row range (its logits land in z[first_row : first_row+n_rows]).
    `lanes` > 1 sizes the workspace for that many chunks in flight on the library's internal streams.
    `src_first_row`: logical row index of patches[0] (a ring buffer that holds only rows [src_first_row, ...) of the
    sequence; the kernels touch rows [first_row, first_row + n_rows) only)."""
    global LAUNCHES
    _chk(patches, torch.float32, 'patches')
    if geo is None:
        total_rows, C, H, W = patches.shape
    else:                                          # `patches` holds whole images (B, C, Himg, Wimg): on-device patchify
        C, (H, W) = patches.shape[1], patch_size
        total_rows = patches.shape[0] * geo.n_per_image
    rows = total_rows - first_row if n_rows is None else n_rows
    lib = _lib.load()
    need = lib.ipsb_resnet_workspace_bytes(desc, chunk, C, H, W) * max(1, int(lanes))
    ws = workspace_cache.get('ws')
    key = (chunk, C, H, W, desc.dt, int(lanes))
    fresh = ws is None or ws.numel() < need or ws.device != patches.device or workspace_cache.get('key') != key
    if fresh:
        if ws is None or ws.numel() < need or ws.device != patches.device:
            ws = torch.empty(need, dtype=torch.uint8, device=patches.device)
        workspace_cache['ws'], workspace_cache['key'] = ws, key
    if z is None:
        z = torch.empty((total_rows, desc.HT), dtype=torch.float32, device=patches.device)
    emb = torch.empty((total_rows, desc.D), dtype=torch.float32, device=patches.device) if want_emb else None
    n_chunks = -(-rows // chunk)
    LAUNCHES += n_chunks * (4 + 2 * desc.n_blocks + sum(int(desc.blocks[i].has_ds) for i in range(desc.n_blocks))
                            + (1 if desc.add_tab else 0))
    z_ptr = z.data_ptr() + first_row * desc.HT * 4
    e_ptr = 0 if emb is None else emb.data_ptr() + first_row * desc.D * 4
    if geo is None:
        src_ptr = _p(patches) - src_first_row * C * H * W * 4
        if src_first_row:
            if not (src_first_row <= first_row and first_row + rows <= src_first_row + patches.shape[0]):
                raise RuntimeError('ips_b200: rows [%d, %d) are not inside the staged window' % (first_row, first_row + rows))
            total_rows = first_row + rows
        _lib.check(lib.ipsb_resnet_logits(desc, src_ptr, first_row, rows, C, H, W, n_per_image, chunk, _p(ws), need,
                                          int(fresh), e_ptr, z_ptr, _stream()))
    else:
        _lib.check(lib.ipsb_resnet_logits_image(desc, _p(patches), ctypes.byref(geo), first_row, rows, C, H, W, chunk, _p(ws), need,
                                                int(fresh), e_ptr, z_ptr, _stream()))
    return z, emb


# ------------------------------------------------------------------ scoring / selection

def score_basis(q_tok, q_w, k_w, H, Dk):
    """U (D, H*T): z[n,h,t] = emb_n . U[:, h*T+t]  (transformer.py:29-31,76-79)."""
    _chk(q_tok, torch.float32, 'q'); _chk(q_w, torch.float32, 'q_w'); _chk(k_w, torch.float32, 'k_w')
    T, D = q_tok.shape[-2], q_tok.shape[-1]
    U = torch.empty((D, H * T), dtype=torch.float32, device=q_tok.device)
    _call('ipsb_score_basis', _p(q_tok), _p(q_w), _p(k_w), _p(U), D, H, Dk, T, _stream())
    return U


def logits(emb, U, add_tab=None, add_idx=None):
    """emb (rows, D) fp32 -> (rows, H*T) logits, plus add_tab[add_idx[row]] when given."""
    _chk(emb, torch.float32, 'emb'); _chk(U, torch.float32, 'U')
    _chk(add_tab, torch.float32, 'add_tab'); _chk(add_idx, torch.int64, 'add_idx')
    rows, D = emb.shape
    HT = U.shape[1]
    z = torch.empty((rows, HT), dtype=torch.float32, device=emb.device)
    _call('ipsb_logits', _p(emb), _p(U), _p(add_tab), _p(add_idx), _p(z), rows, D, HT, _stream())
    return z


def scores_from_logits(z, H, T):
    """z (B,L,H*T) -> scores (B,L)  (transformer.py:143-148)."""
    _chk(z, torch.float32, 'z')
    B, L = z.shape[:2]
    s = torch.empty((B, L), dtype=torch.float32, device=z.device)
    _call('ipsb_scores_from_logits', _p(z), _p(s), B, L, H, T, _stream())
    return s


def topm_stable(scores, M):
    """(values, indices) of the M best per row, ties -> lowest index (ips_net.py:148)."""
    _chk(scores, torch.float32, 'scores')
    B, L = scores.shape
    idx = torch.empty((B, M), dtype=torch.int64, device=scores.device)
    val = torch.empty((B, M), dtype=torch.float32, device=scores.device)
    _call('ipsb_topm_stable', _p(scores), B, L, M, _p(idx), _p(val), _stream())
    return val, idx


def select_loop(z, perm, per_instance, H, T, M, I):
    """Sequential IPS loop on the logit table z (B,N,H*T).  Returns (mem_pos, mem_src, score) (B,M)."""
    _chk(z, torch.float32, 'z'); _chk(perm, torch.int64, 'perm')
    B, N = z.shape[:2]
    dev = z.device
    mem_pos = torch.empty((B, M), dtype=torch.int64, device=dev)
    mem_src = torch.empty((B, M), dtype=torch.int64, device=dev)
    score = torch.empty((B, M), dtype=torch.float32, device=dev)
    need = _lib.load().ipsb_select_loop_workspace_bytes(B, N, H * T, M)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    _call('ipsb_select_loop', _p(z), _p(perm), N if (perm is not None and per_instance) else 0,
                                            B, N, H, T, M, I, _p(mem_pos), _p(mem_src), _p(score), _p(ws), need, _stream())
    return mem_pos, mem_src, score


_SIDE = {}


def _side_stream(dev):
    """One side stream per device for the producer half of the streamed selection."""
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def streamed_select_ok(B, N, HT, M, I):
    """Shapes the streamed selection (`projector_select`) covers: the 8-CTA cluster loop (M + I >= 2048 entries whose
    slices fit shared memory, H*T a power of two)."""
    Lmax = M + min(I, N - M)
    if M >= N or Lmax < 2048 or (HT & (HT - 1)) != 0:
        return False
    cap, Sm = -(-Lmax // 8), -(-M // 8)
    front = max(cap * (HT + 3) * 4, (1 << max(Sm - 1, 0).bit_length()) * 8)
    return front + 8 * Sm * 8 + 12 * 1024 <= 200 * 1024


def projector_select(x, w_bf16, table, B, N, perm, per_instance, H, T, M, I, eps=1e-5, overlap=None):
    """Feature bag (B*N, K) -> the IPS winners, projector and selection loop running SIDE BY SIDE (ips_net.py:200-241: the
    reference embeds chunk i inside iteration i; here the projector kernel reads the rows through the scan order `perm`,
    publishes every 128-row tile of scan-ordered logits with a flag, and the loop's cluster -- launched first, on 8 SMs per
    bag -- consumes the chunks as they complete).  With `overlap=False` (or more than two bags: the loop's clusters would
    take the SMs the projector needs) the two kernels run back to back on the caller's stream; either way no permuted copy
    of the logit table is made.  Returns (mem_pos, mem_src, score, zs, status): zs (B, N, H*T) in scan order, status = the
    loop's two sync words (status[1] != 0: a wait ran out) or None."""
    global LAUNCHES
    _chk(x, x.dtype, 'x'); _chk(w_bf16, torch.bfloat16, 'w'); _chk(table, torch.float32, 'table'); _chk(perm, torch.int64, 'perm')
    HT = H * T
    rows, K = x.shape
    if rows != B * N or not streamed_select_ok(B, N, HT, M, I):
        raise RuntimeError('ips_b200: projector_select does not cover B=%d N=%d H*T=%d M=%d I=%d (see streamed_select_ok)'
                           % (B, N, HT, M, I))
    dev = x.device
    lib = _lib.load()
    if overlap is None:
        overlap = B <= 2 and os.environ.get('IPSB_NO_STREAMED_OVERLAP') is None
    if overlap and not _SIDE.get(('preloaded', dev.index)):
        _lib.check(lib.ipsb_streamed_preload())
        _SIDE[('preloaded', dev.index)] = True
    zs = torch.empty((B, N, HT), dtype=torch.float32, device=dev)
    mem_pos = torch.empty((B, M), dtype=torch.int64, device=dev)
    mem_src = torch.empty((B, M), dtype=torch.int64, device=dev)
    score = torch.empty((B, M), dtype=torch.float32, device=dev)
    need = lib.ipsb_select_loop_workspace_bytes(B, N, HT, M)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    pstride = N if (perm is not None and per_instance) else 0
    is_bf16 = int(x.dtype == torch.bfloat16)
    Nw = w_bf16.shape[0]
    if not overlap:
        _call('ipsb_projector_logits_scan', _p(x), is_bf16, _p(w_bf16), _p(table), _p(zs), rows, K, Nw, HT, eps,
              _p(perm), pstride, N, None, 0, _stream())
        rc = lib.ipsb_select_loop_scan(_p(zs), _p(perm), pstride, B, N, H, T, M, I, _p(mem_pos), _p(mem_src), _p(score),
                                       _p(ws), need, None, 128, None, _stream())
        _lib.check(rc)
        LAUNCHES += 1
        return mem_pos, mem_src, score, zs, None
    tiles = -(-rows // 128)
    words = torch.zeros(tiles + 2, dtype=torch.int32, device=dev)          # [0:2] sync words, [2:] tile flags
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(main)                                                 # flags zeroed, inputs ready
    rc = lib.ipsb_select_loop_scan(_p(zs), _p(perm), pstride, B, N, H, T, M, I, _p(mem_pos), _p(mem_src), _p(score),
                                   _p(ws), need, words.data_ptr() + 8, 128, _p(words), _stream())
    _lib.check(rc)
    with torch.cuda.stream(side):
        _call('ipsb_wait_word', _p(words), _stream())                      # the loop's cluster holds its SMs
        _call('ipsb_projector_logits_scan', _p(x), is_bf16, _p(w_bf16), _p(table), _p(zs), rows, K, Nw, HT, eps,
              _p(perm), pstride, N, words.data_ptr() + 8, torch.cuda.get_device_properties(dev).multi_processor_count - 8 * B, _stream())
    main.wait_stream(side)
    LAUNCHES += 1
    return mem_pos, mem_src, score, zs, words[:2]


def keyed_scan_order(key, rows, N):
    """(rows, N) int64 scan order on the device: row r = the keyed bijection of [0, N) for (key[0], key[1], r); `key` = two
    int64 on the device (ipsb_keyed_scan_order; oracle: ips_oracle.keyed_scan_order)."""
    _chk(key, torch.int64, 'key')
    perm = torch.empty((rows, N), dtype=torch.int64, device=key.device)
    _call('ipsb_keyed_scan_order', _p(key), rows, N, _p(perm), _stream())
    return perm


def merge_candidates(cz, H, T, M):
    """Global re-score of a candidate list cz (B, L, H*T) and stable top-M: positions (B, M) into the list, best first,
    equal scores -> lowest position (score_and_select, ips_net.py:136-155, on the merged buffer).  Long lists run as ONE
    iteration of the cluster selection loop (memory = the first M candidates, chunk = the rest: the same softmax over all
    L entries, the same tie-break), short ones through the single-CTA scores + top-M kernels."""
    B, L = cz.shape[:2]
    if L >= 2048 and ((H * T) & (H * T - 1)) == 0 and L - M >= 1:
        return select_loop(cz, None, False, H, T, M, L - M)[1]
    return topm_stable(scores_from_logits(cz, H, T), M)[1]


def gemm_bf16(mode, a, b, scale=None, shift=None, relu=False, out_dtype=torch.float32):
    """Tensor-core GEMM, bf16 operands, fp32 accumulate.  mode 'nt': a (M,K) b (N,K) -> a b^T;
    'nn': a (M,K) b (K,N) -> a b;  'tn': a (K,M) b (K,N) -> a^T b (split-K, deterministic)."""
    _chk(a, torch.bfloat16, 'a'); _chk(b, torch.bfloat16, 'b'); _chk(scale, torch.float32, 'scale'); _chk(shift, torch.float32, 'shift')
    code = {'nt': 0, 'nn': 1, 'tn': 2}[mode]
    if code == 0:
        (M, K), N = a.shape, b.shape[0]
    elif code == 1:
        (M, K), N = a.shape, b.shape[1]
    else:
        (K, M), N = a.shape, b.shape[1]
    c = torch.empty((M, N), dtype=out_dtype, device=a.device)
    need = _lib.load().ipsb_gemm_workspace_bytes(code, M, N, K)
    ws = torch.empty(need, dtype=torch.uint8, device=a.device) if need else None
    _call('ipsb_gemm_bf16', code, _p(a), _p(b), _p(scale), _p(shift), _p(c), int(out_dtype == torch.float32), M, N, K, int(relu),
          _p(ws), need, _stream())
    return c


def gemm_f32(mode, a, b, shift=None, relu=False):
    """fp32 CUDA-core GEMM in the 'nt' / 'nn' / 'tn' layouts of gemm_bf16."""
    _chk(a, torch.float32, 'a'); _chk(b, torch.float32, 'b'); _chk(shift, torch.float32, 'shift')
    code = {'nt': 0, 'nn': 1, 'tn': 2}[mode]
    if code == 0:
        (M, K), N = a.shape, b.shape[0]
    elif code == 1:
        (M, K), N = a.shape, b.shape[1]
    else:
        (K, M), N = a.shape, b.shape[1]
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _call('ipsb_gemm_f32', code, _p(a), _p(b), _p(shift), _p(c), M, N, K, int(relu), _stream())
    return c


def colsum(x, y=None):
    """sum over rows of x (or of x*y): (rows, cols) -> (cols,)"""
    _chk(x, torch.float32, 'x'); _chk(y, torch.float32, 'y')
    out = torch.empty((x.shape[1],), dtype=torch.float32, device=x.device)
    scratch = torch.empty(64 * x.shape[1], dtype=torch.float32, device=x.device) if x.shape[0] > 4096 else None
    _call('ipsb_colsum_f32', _p(x), _p(y), _p(out), _p(scratch), x.shape[0], x.shape[1], _stream())
    return out


def cast_bf16(x):
    _chk(x, torch.float32, 'x')
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _call('ipsb_cast_bf16', _p(x), _p(y), x.numel(), _stream())
    return y


# ------------------------------------------------------------------ aggregator + heads (no-grad forward)

def cross_attention(q_scaled, k, v, H, Dk, Dv):
    """q_scaled (T,H*Dk), k (B,M,H*Dk), v (B,M,H*Dv) -> (B,T,H*Dv): softmax(q k^T) v per head."""
    _chk(q_scaled, torch.float32, 'q'); _chk(k, torch.float32, 'k'); _chk(v, torch.float32, 'v')
    B, M = k.shape[:2]
    T = q_scaled.shape[0]
    out = torch.empty((B, T, H * Dv), dtype=torch.float32, device=k.device)
    _call('ipsb_cross_attention_f32', _p(q_scaled), _p(k), _p(v), _p(out), B, M, H, Dk, Dv, T, _stream())
    return out


def residual_layernorm(x, r, gamma, beta, eps):
    """LayerNorm(x + r) * gamma + beta; x (rows, D), r (r_rows, D) broadcast over rows or None."""
    _chk(x, torch.float32, 'x'); _chk(r, torch.float32, 'r'); _chk(gamma, torch.float32, 'gamma'); _chk(beta, torch.float32, 'beta')
    rows, D = x.shape
    y = torch.empty_like(x)
    _call('ipsb_residual_layernorm_f32', _p(x), _p(r), 0 if r is None else r.shape[0], _p(gamma), _p(beta), _p(y), rows, D,
          eps, _stream())
    return y


def head_activation(logits, act):
    _chk(logits, torch.float32, 'logits')
    y = torch.empty_like(logits)
    _call('ipsb_head_activation_f32', _p(logits), _p(y), logits.shape[0], logits.shape[1], 1 if act == 'sigmoid' else 0, _stream())
    return y


def add(a, b):
    _chk(a, torch.float32, 'a'); _chk(b, torch.float32, 'b')
    y = torch.empty_like(a)
    _call('ipsb_add_f32', _p(a), _p(b), _p(y), a.numel(), _stream())
    return y


# ------------------------------------------------------------------ per-launch timing of the native executor

PROFILE_KINDS = ('stage', 'stem', 'conv', 'pool', 'logits')


def profile_begin():
    """Start bracketing every kernel the native executor issues with CUDA events (measurement only)."""
    _lib.check(_lib.load().ipsb_profile_begin(_stream()))


def profile_end(max_records=1 << 16):
    """Stop; returns [(family, start_ms, stop_ms)] relative to profile_begin, after synchronising the events."""
    kinds = (ctypes.c_int32 * max_records)()
    t0 = (ctypes.c_float * max_records)()
    t1 = (ctypes.c_float * max_records)()
    n = ctypes.c_int32(0)
    _lib.check(_lib.load().ipsb_profile_end(max_records, ctypes.addressof(kinds), ctypes.addressof(t0), ctypes.addressof(t1),
                                            ctypes.byref(n)))
    return [(PROFILE_KINDS[kinds[i]], t0[i], t1[i]) for i in range(n.value)]


def busy_ms(intervals):
    """Length of the union of (start, stop) intervals."""
    total, cur_a, cur_b = 0.0, None, None
    for a, b in sorted(intervals):
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                total += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        total += cur_b - cur_a
    return total


# ------------------------------------------------------------------ NVLink peer exchange (sequence-sharded selection)

PEER_HEADER = 4096
_peer_maps = {}          # (device index, IPC handle bytes) -> base address of the mapping in this process


def peer_export(buf):
    """(handle bytes[64], offset) of a CUDA tensor's storage for ipsb_peer_open in another process."""
    handle = (ctypes.c_ubyte * 64)()
    off = ctypes.c_int64(0)
    _lib.check(_lib.load().ipsb_peer_export(buf.data_ptr(), ctypes.addressof(handle), ctypes.byref(off)))
    return bytes(handle), off.value


def peer_open(handle, offset, device_index):
    """Address in this process of a peer's exchange buffer; one mapping per peer allocation is kept for the process."""
    key = (device_index, handle)
    if key not in _peer_maps:
        raw = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        out = ctypes.c_void_p(0)
        _lib.check(_lib.load().ipsb_peer_open(ctypes.addressof(raw), 0, ctypes.byref(out)))
        _peer_maps[key] = out.value
    return _peer_maps[key] + offset


def peer_ctx(rank, world, bases):
    c = _lib.PeerCtx()
    c.rank, c.world = rank, world
    for q, b in enumerate(bases):
        c.base[q] = b
    return c


def peer_push_candidates(ctx, z_local, cand, index_base, L, slot0, cz_off, ci_off, rows=None):
    """`rows` (B, m): rows of z_local holding the candidates' logits when the table is in scan order (default: cand)."""
    _chk(z_local, torch.float32, 'z_local'); _chk(cand, torch.int64, 'cand'); _chk(rows, torch.int64, 'rows')
    B, n_local, HT = z_local.shape
    _call('ipsb_peer_push_candidates_rows', ctypes.byref(ctx), _p(z_local), n_local, _p(cand), _p(rows), B, cand.shape[1], HT,
          index_base, L, slot0, cz_off, ci_off, _stream())


def peer_push_logits(ctx, z_local, N, row0, z_off):
    _chk(z_local, torch.float32, 'z_local')
    B, n_local, HT = z_local.shape
    _call('ipsb_peer_push_logits', ctypes.byref(ctx), _p(z_local), B, n_local, HT, N, row0, z_off, _stream())


def peer_allgather_small(ctx, src, sec_off, slot_stride, phase, world):
    """src (contiguous fp32 CUDA vector, length a multiple of 4) -> (world, len) tensor with every rank's vector."""
    _chk(src, torch.float32, 'src')
    out = torch.empty((world, src.numel()), dtype=torch.float32, device=src.device)
    _call('ipsb_peer_allgather_small', ctypes.byref(ctx), _p(src), src.numel() * 4, sec_off, slot_stride, phase, _p(out), _stream())
    return out


def peer_bn_forward(ctx, mean_var, cols, sec_off, slot_stride, phase, momentum, unbias, eps, running_mean, running_var):
    """Synchronised BatchNorm forward exchange: every rank's [mean | var] -> (mean, rstd) of the whole batch, running
    statistics updated in the waiting kernel (ipsb_peer_bn_forward)."""
    _chk(mean_var, torch.float32, 'mean_var'); _chk(running_mean, torch.float32, 'running_mean'); _chk(running_var, torch.float32, 'running_var')
    mean = torch.empty(cols, dtype=torch.float32, device=mean_var.device)
    rstd = torch.empty_like(mean)
    _call('ipsb_peer_bn_forward', ctypes.byref(ctx), _p(mean_var), cols, sec_off, slot_stride, phase, float(momentum), float(unbias),
          float(eps), _p(mean), _p(rstd), _p(running_mean), _p(running_var), _stream())
    return mean, rstd


def peer_allgather_sum(ctx, src, sec_off, slot_stride, phase):
    """Sum over the ranks (rank order) of `src` (contiguous fp32, length a multiple of 4)."""
    _chk(src, torch.float32, 'src')
    out = torch.empty_like(src)
    _call('ipsb_peer_allgather_sum', ctypes.byref(ctx), _p(src), src.numel(), sec_off, slot_stride, phase, _p(out), _stream())
    return out


def peer_wait(ctx, phase):
    _call('ipsb_peer_wait', ctypes.byref(ctx), phase, _stream())


def peer_push_winners(ctx, local_rows, row0, win, ci, L, row_bytes, slides_per_rank, out_off, idx_out):
    _chk(win, torch.int64, 'win'); _chk(ci, torch.int64, 'ci'); _chk(idx_out, torch.int64, 'idx_out')
    if not local_rows.is_cuda or not local_rows.is_contiguous():
        raise RuntimeError('ips_b200: local patches must be a contiguous CUDA tensor')
    B, M = win.shape
    _call('ipsb_peer_push_winners', ctypes.byref(ctx), _p(local_rows), local_rows.shape[1], row0, _p(win), _p(ci), L, B, M, row_bytes,
          slides_per_rank, out_off, _p(idx_out), _stream())


def peer_status(ctx):
    out = ctypes.c_int32(0)
    _lib.check(_lib.load().ipsb_peer_status(ctypes.byref(ctx), ctypes.byref(out), _stream()))
    return out.value


# ------------------------------------------------------------------ torch.ops registration
# Thin public aliases so the kernels are reachable as torch.ops.ips_b200.*; the
# module code calls the Python functions above directly (lower dispatch overhead).
_registered = False


def register_custom_ops():
    global _registered
    if _registered:
        return
    lib = torch.library.Library('ips_b200', 'DEF')
    lib.define('scores_from_logits(Tensor z, int H, int T) -> Tensor')
    lib.define('topm_stable(Tensor scores, int M) -> (Tensor, Tensor)')
    lib.define('select_loop(Tensor z, Tensor? perm, bool per_instance, int H, int T, int M, int I) -> (Tensor, Tensor, Tensor)')
    lib.define('gather_rows(Tensor src, Tensor idx, int batch_stride_rows) -> Tensor')
    lib.define('logits(Tensor emb, Tensor U, Tensor? add_tab, Tensor? add_idx) -> Tensor')
    lib.impl('scores_from_logits', scores_from_logits, 'CUDA')
    lib.impl('topm_stable', topm_stable, 'CUDA')
    lib.impl('select_loop', select_loop, 'CUDA')
    lib.impl('gather_rows', gather_rows, 'CUDA')
    lib.impl('logits', logits, 'CUDA')
    register_custom_ops._lib = lib      # keep alive
    _registered = True
