"""ctypes binding of the C ABI declared in include/ips_b200.h.

There is no CPU or PyTorch fallback: if libips_b200.so is missing (or was built
for another architecture) every op raises.  Build it with
``python -m ips_b200.build`` or ``__graft_entry__.build()``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libips_b200.so')

_i32, _i64, _f32, _ptr = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p



class ConvDesc(ctypes.Structure):
    """ipsb_conv_desc"""
    _fields_ = [('w', _ptr), ('scale', _ptr), ('shift', _ptr),
                ('cin', ctypes.c_int32), ('cout', ctypes.c_int32), ('kh', ctypes.c_int32), ('kw', ctypes.c_int32),
                ('stride', ctypes.c_int32), ('pad', ctypes.c_int32), ('mode', ctypes.c_int32), ('_pad', ctypes.c_int32)]


class BlockDesc(ctypes.Structure):
    """ipsb_block_desc"""
    _fields_ = [('c1', ConvDesc), ('c2', ConvDesc), ('ds', ConvDesc), ('has_ds', ctypes.c_int32), ('_pad', ctypes.c_int32)]


class ResnetDesc(ctypes.Structure):
    """ipsb_resnet_desc"""
    _fields_ = [('dt', ctypes.c_int32), ('n_blocks', ctypes.c_int32), ('stem', ConvDesc), ('blocks', BlockDesc * 8),
                ('D', ctypes.c_int32), ('HT', ctypes.c_int32), ('U', _ptr), ('add_tab', _ptr)]


class ImageGeo(ctypes.Structure):
    """ipsb_image_geo"""
    _fields_ = [('img_h', ctypes.c_int32), ('img_w', ctypes.c_int32), ('stride_h', ctypes.c_int32), ('stride_w', ctypes.c_int32),
                ('n_per_image', ctypes.c_int32)]


class FoldItem(ctypes.Structure):
    """ipsb_fold_item"""
    _fields_ = [('w_src', _ptr), ('w_dst', _ptr), ('bn_weight', _ptr), ('bn_bias', _ptr), ('bn_mean', _ptr), ('bn_var', _ptr),
                ('scale_dst', _ptr), ('shift_dst', _ptr), ('dst_elems', ctypes.c_int64),
                ('cout', ctypes.c_int32), ('cin', ctypes.c_int32), ('kh', ctypes.c_int32), ('kw', ctypes.c_int32),
                ('cin_pad', ctypes.c_int32), ('layout', ctypes.c_int32), ('dst_bf16', ctypes.c_int32), ('eps', ctypes.c_float)]


FOLD_KMAJOR, FOLD_KN, FOLD_STEM_S2D, FOLD_STEM_8X8 = 0, 1, 2, 3


class PeerCtx(ctypes.Structure):
    """ipsb_peer_ctx"""
    _fields_ = [('rank', ctypes.c_int32), ('world', ctypes.c_int32), ('base', _ptr * 8)]


# name -> argtypes (restype is int status unless listed in _RESTYPE)
SIGNATURES = {
    'ipsb_abi_version': [],
    'ipsb_last_error': [],
    'ipsb_device_ok': [],
    'ipsb_stage_patches': [_ptr, _ptr, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_stage_patches_padded': [_ptr, _ptr, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_stage_patches_s2d': [_ptr, _ptr, _i64, _i64, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_stage_patches_s2d_tma': [_ptr, _ptr, _i64, _i64, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_stage_tma_ok': [_ptr, _i32, _i32, _i32],
    'ipsb_stage_image_s2d': [_ptr, ctypes.POINTER(ImageGeo), _i64, _i64, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_gather_patches_image': [_ptr, ctypes.POINTER(ImageGeo), _ptr, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_im2col_bf16': [_ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_maxpool3x3s2_pf_strided': [_ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_stem_pool_s2d': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_conv_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_linear_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_maxpool3x3s2': [_ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_avgpool': [_ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_layernorm_rows_f32': [_ptr, _ptr, _i64, _i32, _f32, _ptr],
    'ipsb_conv_bf16_umma': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_pf_rows': [_i64, _i32, _i32],
    'ipsb_conv_bf16_pf': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_maxpool3x3s2_pf': [_ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_avgpool_pf': [_ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_linear_bf16_umma': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_gemm_workspace_bytes': [_i32, _i64, _i32, _i64],
    'ipsb_gemm_bf16': [_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i64, _i32, _i64, _i32, _ptr, _i64, _ptr],
    'ipsb_gemm_f32': [_i32, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i64, _i32, _ptr],
    'ipsb_colsum_f32': [_ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr],
    'ipsb_cast_bf16': [_ptr, _ptr, _i64, _ptr],
    'ipsb_rows_to_bf16': [_ptr, _ptr, _i64, _i32, _i32, _f32, _ptr],
    'ipsb_rows_bf16_to_bf16': [_ptr, _ptr, _i64, _i32, _i32, _f32, _ptr],
    'ipsb_score_basis': [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_logits': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr],
    'ipsb_scores_from_logits': [_ptr, _ptr, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_topm_stable': [_ptr, _i32, _i32, _i32, _ptr, _ptr, _ptr],
    'ipsb_select_loop_workspace_bytes': [_i32, _i32, _i32, _i32],
    'ipsb_select_loop': [_ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _i64, _ptr],
    'ipsb_cross_attention_f32': [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_residual_layernorm_f32': [_ptr, _ptr, _i32, _ptr, _ptr, _ptr, _i64, _i32, _f32, _ptr],
    'ipsb_head_loss_f32': [_ptr, _ptr, _ptr, _i32, _i32, _i32, _f32, _ptr, _ptr, _ptr, _ptr],
    'ipsb_head_activation_f32': [_ptr, _ptr, _i32, _i32, _i32, _ptr],
    'ipsb_add_f32': [_ptr, _ptr, _ptr, _i64, _ptr],
    'ipsb_bn_stats_f32': [_ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr],
    'ipsb_bn_stats_finalize_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _f32, _f32, _f32, _ptr, _i64, _i32, _ptr],
    'ipsb_maxpool3x3s2_bwd_f32': [_ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_add_relu_f32': [_ptr, _ptr, _ptr, _i64, _ptr],
    'ipsb_relu_bwd_f32': [_ptr, _ptr, _ptr, _i64, _ptr],
    'ipsb_wgrad_to_oihw': [_ptr, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_conv_weight_layouts': [_ptr, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr],
    'ipsb_bn_finalize_f32': [_ptr, _ptr, _i32, _f32, _f32, _f32, _ptr, _ptr, _ptr, _ptr],
    'ipsb_bn_apply_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr],
    'ipsb_bn_backward_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr],
    'ipsb_bn_backward_sums_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr],
    'ipsb_bn_backward_apply_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i64, _i32, _ptr],
    'ipsb_layernorm_backward_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _f32, _ptr],
    'ipsb_attention_chunks': [_i32],
    'ipsb_attention_train_fwd_f32': [_ptr, _ptr, _ptr, _ptr, _f32, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_attention_train_bwd_f32': [_ptr, _ptr, _ptr, _ptr, _f32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _i32,
                                     _ptr],
    'ipsb_gather_rows': [_ptr, _i64, _ptr, _i32, _i32, _i64, _ptr, _ptr],
    'ipsb_conv_bf16_f32out': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_sum3_split': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _i64, _i32, _ptr],
    'ipsb_sum3_maxpool_split': [_ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr, _ptr, _i32, _ptr],
    'ipsb_stage_patches_padded_split': [_ptr, _i64, _i64, _i32, _i32, _i32, _ptr, _ptr, _ptr],
    'ipsb_fold_plan': [_ptr, _i32, _i32, _ptr],
    'ipsb_projector_logits': [_ptr, _i32, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _f32, _ptr],
    'ipsb_projector_logits_scan': [_ptr, _i32, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _f32, _ptr, _i64, _i64, _ptr, _i32, _ptr],
    'ipsb_select_loop_scan': [_ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i32, _ptr, _ptr],
    'ipsb_wait_word': [_ptr, _ptr],
    'ipsb_keyed_scan_order': [_ptr, _i32, _i32, _ptr, _ptr],
    'ipsb_streamed_preload': [],
    'ipsb_projector_preload': [],
    'ipsb_profile_begin': [_ptr],
    'ipsb_profile_end': [_i32, _ptr, _ptr, _ptr, ctypes.POINTER(_i32)],
    'ipsb_peer_export': [_ptr, _ptr, ctypes.POINTER(_i64)],
    'ipsb_peer_open': [_ptr, _i64, ctypes.POINTER(_ptr)],
    'ipsb_peer_close': [_ptr, _i64],
    'ipsb_peer_status': [ctypes.POINTER(PeerCtx), ctypes.POINTER(_i32), _ptr],
    'ipsb_peer_push_candidates_rows': [ctypes.POINTER(PeerCtx), _ptr, _i64, _ptr, _ptr, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i64, _ptr],
    'ipsb_peer_push_candidates': [ctypes.POINTER(PeerCtx), _ptr, _i64, _ptr, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i64, _ptr],
    'ipsb_peer_push_logits': [ctypes.POINTER(PeerCtx), _ptr, _i32, _i64, _i32, _i64, _i64, _i64, _ptr],
    'ipsb_peer_bn_forward': [ctypes.POINTER(PeerCtx), _ptr, _i32, _i64, _i64, _i32, _f32, _f32, _f32, _ptr, _ptr, _ptr, _ptr, _ptr],
    'ipsb_peer_allgather_sum': [ctypes.POINTER(PeerCtx), _ptr, _i32, _i64, _i64, _i32, _ptr, _ptr],
    'ipsb_peer_allgather_small': [ctypes.POINTER(PeerCtx), _ptr, _i64, _i64, _i64, _i32, _ptr, _ptr],
    'ipsb_peer_wait': [ctypes.POINTER(PeerCtx), _i32, _ptr],
    'ipsb_peer_push_winners': [ctypes.POINTER(PeerCtx), _ptr, _i64, _i64, _ptr, _ptr, _i64, _i32, _i32, _i64, _i32, _i64, _ptr, _ptr],
    'ipsb_resnet_workspace_bytes': [ctypes.POINTER(ResnetDesc), _i64, _i32, _i32, _i32],
    'ipsb_resnet_logits': [ctypes.POINTER(ResnetDesc), _ptr, _i64, _i64, _i32, _i32, _i32, _i64, _i64, _ptr, _i64, _i32,
                           _ptr, _ptr, _ptr],
    'ipsb_resnet_logits_image': [ctypes.POINTER(ResnetDesc), _ptr, ctypes.POINTER(ImageGeo), _i64, _i64, _i32, _i32, _i32, _i64,
                                 _ptr, _i64, _i32, _ptr, _ptr, _ptr],
}
_RESTYPE = {'ipsb_last_error': ctypes.c_char_p, 'ipsb_resnet_workspace_bytes': ctypes.c_int64, 'ipsb_attention_chunks': ctypes.c_int64,
            'ipsb_pf_rows': ctypes.c_int64, 'ipsb_select_loop_workspace_bytes': ctypes.c_int64,
            'ipsb_gemm_workspace_bytes': ctypes.c_int64}

_lib = None


def load():
    """Load the shared library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: the CUDA library is the only implementation of this package. '
            'Build it with `python -m ips_b200.build` (needs nvcc with sm_100a support).')
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, ctypes.c_int)
    if lib.ipsb_abi_version() != 1:
        raise RuntimeError('libips_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise RuntimeError('ips_b200: ' + load().ipsb_last_error().decode())
