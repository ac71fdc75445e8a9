"""ctypes binding of the C ABI in include/lft_b200.h (the only way Python reaches the kernels).

There is deliberately no fallback: if the shared library is missing it is built with nvcc; if that
is impossible, or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)


class LftConfig(C.Structure):
    _fields_ = [("ang_res", C.c_int32), ("scale", C.c_int32), ("channels", C.c_int32),
                ("precision", C.c_int32), ("device", C.c_int32)]


class LftError(RuntimeError):
    pass


PREC_FP32 = 0
PREC_BF16 = 1
PROFILE_MAX_KINDS = 16

# name -> (restype, argtypes); mirrors include/lft_b200.h one to one
SIGNATURES = {
    "lft_last_error": (C.c_char_p, []),
    "lft_version": (C.c_int, []),
    "lft_create": (C.c_int, [C.POINTER(LftConfig), C.POINTER(C.c_void_p)]),
    "lft_destroy": (C.c_int, [C.c_void_p]),
    "lft_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_i64_p, C.c_int32]),
    "lft_finalize_weights": (C.c_int, [C.c_void_p]),
    "lft_set_precision": (C.c_int, [C.c_void_p, C.c_int32]),
    "lft_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "lft_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                              C.c_void_p]),
    "lft_lf_num_patches": (C.c_int, [C.c_int32, C.c_int32, c_i32_p, c_i32_p]),
    "lft_forward_lf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "lft_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p]),
    "lft_divide": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                             C.c_void_p]),
    "lft_lf_num_patches_ex": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i32_p, c_i32_p]),
    "lft_forward_lf_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lft_forward_lf_sr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lft_peer_alloc": (C.c_int, [C.c_int32, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "lft_peer_free": (C.c_int, [C.c_int32, C.c_void_p]),
    "lft_peer_open": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "lft_peer_close": (C.c_int, [C.c_int32, C.c_void_p]),
    "lft_integrate_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p, C.c_void_p]),
    "lft_divide_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_void_p, C.c_void_p]),
    "lft_stage_conv_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    "lft_stage_ang": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    "lft_stage_spa": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    "lft_stage_upsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "lft_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "lft_profile_read": (C.c_int, [C.c_void_p, c_i32_p, C.POINTER(C.c_char_p), c_i64_p, C.POINTER(C.c_double)]),
    "lft_profile_read2": (C.c_int, [C.c_void_p, c_i32_p, C.POINTER(C.c_char_p), c_i64_p, C.POINTER(C.c_double), c_i64_p]),
    "lft_launch_count": (C.c_int64, [C.c_void_p]),
    "lft_debug_timeline": (C.c_int, [C.c_int32, c_i64_p]),
    "lft_mma_bench": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i64_p]),
    "lft_gemm_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32]),
}


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load liblft_b200.so (building it in-tree first if needed). Raises if it cannot be had."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        if not os.path.exists(path):
            if not build_if_missing:
                raise LftError(f"{path} not built; run `python -m lft_b200.build`")
            _build.build()
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().lft_last_error()
        raise LftError(f"lft_b200 error {rc}: {msg.decode() if msg else '?'}")
