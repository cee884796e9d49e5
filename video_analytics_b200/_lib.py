"""ctypes binding of libva_b200.so -- the ONLY compute backend of this package.

There is no CPU or PyTorch fallback: if the shared library is missing, or a call fails, an exception is raised.
Signatures mirror include/va_b200.h one to one.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libva_b200.so"


class VAError(RuntimeError):
    """A libva_b200 call returned non-zero; message comes from va_last_error()."""


# every symbol include/va_b200.h declares: name -> (restype, argtypes)
_vp, _i, _f, _sz, _u32 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_uint32
SIGNATURES = {
    "va_last_error": (C.c_char_p, []),
    "va_abi_version": (_i, []),
    "va_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "va_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "va_create_ex": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "va_destroy": (_i, [_vp]),
    "va_input_channels_padded": (_i, [_vp]),
    "va_load_weights": (_i, [_vp, C.POINTER(_vp), _i, _vp]),
    "va_preprocess": (_i, [_vp, _sz, _i, _i, _i, _vp, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _i, _i, _vp, _vp]),
    "va_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "va_forward_store": (_i, [_vp, _vp, _sz, _i, _i, _i, _vp, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp, _vp, _vp, _vp]),
    "va_conv1_fused": (_i, [_vp, _sz, _i, _i, _i, _vp, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp, _vp, _vp]),
    "va_conv2d_nhwc": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "va_linear": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "va_fuse": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "va_svm_decision": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "va_consensus_update": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "va_pack_input_nchw": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "va_pack_input_nchw_split6": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "va_maxpool2x2_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "va_pool_bwd_codes": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "va_relu_pool_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "va_bias_grad": (_i, [_vp, C.c_longlong, _i, _vp, _vp]),
    "va_dropout": (_i, [_vp, _vp, C.c_longlong, _f, _i, _vp, _vp]),
    "va_conv2d_dgrad": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "va_linear_dgrad": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "va_wgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "va_ce_train": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "va_relu_bwd_f32_to_bf16": (_i, [_vp, _vp, C.c_longlong, _vp, _vp]),
    "va_sgd_momentum": (_i, [_vp, _vp, _vp, C.c_longlong, _f, _f, _i, _f, _vp]),
    "va_sgd_momentum_bf16g": (_i, [_vp, _vp, _vp, C.c_longlong, _f, _f, _i, _f, _vp]),
    "va_transpose_bf16": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "va_f32_to_bf16": (_i, [_vp, C.c_longlong, _vp, _vp]),
    "va_jpeg_decode": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "va_svm_fit": (_i, [_vp, _vp, _i, _i, _i, C.c_double, C.c_double, C.c_double, _i, _vp, _vp, _vp, _vp, _vp]),
    "va_tvl1_workspace_bytes": (_sz, [_i, _i, _vp]),
    "va_tvl1_flow": (_i, [_vp, _sz, _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "va_tvl1_debug_cycles": (_i, [_vp]),
    "va_reserve_sms": (_i, [_i, _i]),
    "va_allreduce_bf16": (_i, [_vp, _vp, _i, _i, C.c_longlong, _i, _vp]),
    "va_synth_fill": (_i, [_vp, _sz, _i, _i, _i, _i, _u32, _u32, _vp]),
    "va_debug_conv_counters": (_i, [_vp]),
    "va_profile_enable": (_i, [_i]),
    "va_profile_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "va_launch_count": (C.c_uint64, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libva_b200.so (built in-tree by video_analytics_b200.build) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VAError(
            f"{LIB_PATH} is missing: build it with `python -m video_analytics_b200.build` "
            "(there is no CPU/PyTorch fallback for the two-stream path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().va_last_error()
        raise VAError(f"{what or 'libva_b200'} failed ({status}): {msg.decode() if msg else '?'}")


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None) -> C.c_void_p:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def launch_count() -> int:
    return int(load().va_launch_count())
