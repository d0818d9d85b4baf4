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

# name -> argtypes (restype is int status unless listed in _RESTYPE)
SIGNATURES = {
    'ipsb_abi_version': [],
    'ipsb_last_error': [],
    'ipsb_device_ok': [],
    'ipsb_stage_patches': [_ptr, _ptr, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr],
    'ipsb_conv_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_linear_f32': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_maxpool3x3s2': [_ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_avgpool': [_ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_layernorm_rows_f32': [_ptr, _ptr, _i64, _i32, _f32, _ptr],
    'ipsb_conv_bf16_umma': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_linear_bf16_umma': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _ptr],
    'ipsb_rows_to_bf16': [_ptr, _ptr, _i64, _i32, _i32, _f32, _ptr],
    'ipsb_score_basis': [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_logits': [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr],
    'ipsb_scores_from_logits': [_ptr, _ptr, _i32, _i32, _i32, _i32, _ptr],
    'ipsb_topm_stable': [_ptr, _i32, _i32, _i32, _ptr, _ptr, _ptr],
    'ipsb_select_loop': [_ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr],
    'ipsb_gather_rows': [_ptr, _i64, _ptr, _i32, _i32, _i64, _ptr, _ptr],
}
_RESTYPE = {'ipsb_last_error': ctypes.c_char_p}

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
