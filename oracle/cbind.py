"""ctypes access to oracle/_c/liboracle.so (C restatement; test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .cfft import layer_twiddles
from .circle import CanonicCoset
from .fields import P, U64, m_inv

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_c", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _tw_ptrs(tws):
    arrs = [np.ascontiguousarray(t, dtype=np.uint32) for t in tws]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    return arrs, ptrs


class CpuCfft:
    """Pre-built twiddles for one log size; evaluate / interpolate n_cols columns in place."""

    def __init__(self, log_n: int):
        self.log_n = log_n
        dom = CanonicCoset(log_n).circle_domain()
        tws = layer_twiddles(dom)
        self._fw, self._fwp = _tw_ptrs(tws)
        self._iw, self._iwp = _tw_ptrs([m_inv(t) for t in tws])
        self.inv_n = pow(1 << log_n, P - 2, P)

    def evaluate(self, v: np.ndarray, n_threads: int = 0):
        assert v.dtype == np.uint32 and v.flags.c_contiguous and v.shape[-1] == 1 << self.log_n
        n_cols = v.size >> self.log_n
        lib().oracle_cfft_evaluate(v.ctypes.data_as(C.c_void_p), self.log_n, n_cols, self._fwp, n_threads)

    def interpolate(self, v: np.ndarray, n_threads: int = 0):
        assert v.dtype == np.uint32 and v.flags.c_contiguous and v.shape[-1] == 1 << self.log_n
        n_cols = v.size >> self.log_n
        lib().oracle_cfft_interpolate(v.ctypes.data_as(C.c_void_p), self.log_n, n_cols, self._iwp, C.c_uint32(self.inv_n), n_threads)


def merkle_layer(prev, cols, n_nodes: int) -> np.ndarray:
    """prev: uint32[2*n_nodes*8] or None; cols: list of uint32 arrays (>= n_nodes)."""
    out = np.empty(n_nodes * 8, dtype=np.uint32)
    cols = [np.ascontiguousarray(c, dtype=np.uint32) for c in cols]
    ptrs = (C.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
    pp = prev.ctypes.data_as(C.c_void_p) if prev is not None else None
    lib().oracle_merkle_layer(out.ctypes.data_as(C.c_void_p), pp, ptrs, len(cols), C.c_size_t(n_nodes))
    return out


def max_threads() -> int:
    return lib().oracle_max_threads()
