"""ctypes binding of include/luminair_b200.h (no torch types cross this boundary)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libluminair_b200.so")

LB_OK = 0
_ERRS = {-1: "LB_ERR_CUDA", -2: "LB_ERR_OOM", -3: "LB_ERR_BAD_ARG", -4: "LB_ERR_NCCL"}


class LuminairB200Error(RuntimeError):
    pass


_lib = None

u32p = C.POINTER(C.c_uint32)
ctxp = C.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/luminair_b200.h
SIGNATURES = {
    "lb_version": (C.c_int, []),
    "lb_ctx_create": (C.c_int, [C.c_int, C.POINTER(ctxp)]),
    "lb_ctx_destroy": (None, [ctxp]),
    "lb_last_error": (C.c_char_p, [ctxp]),
    "lb_sync": (C.c_int, [ctxp]),
    "lb_alloc": (C.c_int, [ctxp, C.c_size_t, C.POINTER(C.c_void_p)]),
    "lb_free": (C.c_int, [ctxp, C.c_void_p]),
    "lb_memset_zero": (C.c_int, [ctxp, C.c_void_p, C.c_size_t]),
    "lb_upload": (C.c_int, [ctxp, C.c_void_p, C.c_void_p, C.c_size_t]),
    "lb_download": (C.c_int, [ctxp, C.c_void_p, C.c_void_p, C.c_size_t]),
    "lb_copy": (C.c_int, [ctxp, C.c_void_p, C.c_void_p, C.c_size_t]),
    "lb_timer_start": (C.c_int, [ctxp]),
    "lb_timer_stop_ms": (C.c_int, [ctxp, C.POINTER(C.c_float)]),
    "lb_device_info": (C.c_int, [ctxp, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "lb_twiddles_ensure": (C.c_int, [ctxp, C.c_int]),
    "lb_twiddles_export": (C.c_int, [ctxp, C.c_int, C.c_void_p]),
    "lb_interpolate_batch": (C.c_int, [ctxp, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "lb_evaluate_batch": (C.c_int, [ctxp, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "lb_merkle_commit_layer": (C.c_int, [ctxp, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "lb_gather_rows": (C.c_int, [ctxp, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
}


def load_library():
    """Load libluminair_b200.so; fail loudly when it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LuminairB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). luminair_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc: int, what: str = ""):
    if rc != LB_OK:
        msg = load_library().lb_last_error(ctx)
        raise LuminairB200Error(f"{what}: {_ERRS.get(rc, rc)}: {msg.decode() if msg else ''}")
