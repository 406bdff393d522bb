"""ctypes binding of ``libl2i_b200.so`` (the C ABI declared in ``include/l2i_b200.h``).

PyTorch is used here only for device memory and streams: every call passes raw
``tensor.data_ptr()`` values and the current CUDA stream handle.  There is no CPU or PyTorch
fallback: if the library is missing, or a kernel reports an error, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

F32, BF16, F16 = 0, 1, 2
_DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

# L2I_LIB: debug override (A/B runs of differently compiled builds); the default is the in-tree build
_LIB_PATH = os.environ.get("L2I_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libl2i_b200.so")
_lib: Optional[C.CDLL] = None

_vp, _i64, _i32, _f32, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64

# name -> (restype, argtypes); mirrors include/l2i_b200.h one to one
_SIGNATURES = {
    "l2i_abi_version": (_i32, []),
    "l2i_last_error_string": (C.c_char_p, []),
    "l2i_launch_count": (_i64, []),
    "l2i_fused_bias_act": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _f32, _f32, _i32, _vp]),
    "l2i_fused_leaky_relu_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _f32, _i32, _vp]),
    "l2i_upfirdn2d": (_i32, [_vp, _vp, _vp, _i64] + [_i32] * 13 + [_i32, _vp]),
    "l2i_linear_fwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _i32, _f32, _f32, _vp]),
    "l2i_pixel_norm": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "l2i_walk_linear_fwd": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _u64, _vp]),
    "l2i_walk_linear_bwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _u64, _vp]),
    "l2i_walk_combine": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i32, _i32, _i32, _u64, _i32, _vp]),
    "l2i_linear_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "l2i_walk_combine_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _i32, _i32, _u64, _i32, _vp]),
    "l2i_generator_create": (_i32, [C.POINTER(_vp), _i32, _i32, _i32, _i32, C.POINTER(_f32), _i32, _f32, _i32, _i32]),
    "l2i_generator_destroy": (None, [_vp]),
    "l2i_generator_set_param": (_i32, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "l2i_generator_finalize": (_i32, [_vp, _vp]),
    "l2i_generator_num_layers": (_i32, [_vp]),
    "l2i_generator_n_latent": (_i32, [_vp]),
    "l2i_generator_mapping": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "l2i_generator_forward": (_i32, [_vp, _vp, _i64, _i64, C.POINTER(_vp), C.POINTER(_i32), _vp, _vp, _i32, _vp]),
    "l2i_generator_set_profiling": (_i32, [_vp, _i32]),
    "l2i_generator_profile_count": (_i32, [_vp]),
    "l2i_generator_profile_entry": (_i32, [_vp, _i32, C.c_char_p, _i32, C.POINTER(_i32), C.POINTER(_f32),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "l2i_generator_set_training": (_i32, [_vp, _i32]),
    "l2i_generator_backward": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "l2i_generator_read_activation": (_i32, [_vp, C.c_char_p, _vp, _i64, _i32, _vp]),
    "l2i_image_to_uint8": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Loads the shared library once; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"latent2im_b200: native library not found at {_LIB_PATH}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().l2i_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"latent2im_b200: {what} failed (code {rc}): {msg}")


def dtype_code(t: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[t]
    except KeyError:
        raise RuntimeError(f"latent2im_b200: unsupported dtype {t} (float32 / bfloat16 / float16 only)")


def require_cuda(t: torch.Tensor, name: str) -> None:
    # same message as the reference's CHECK_CUDA (op/upfirdn2d.cpp:8, op/fused_bias_act.cpp:7)
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def launch_count() -> int:
    return int(load().l2i_launch_count())
