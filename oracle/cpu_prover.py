"""ctypes access to the compiled CPU prover oracle/c/cpu_prover (oracle; TEST INFRASTRUCTURE ONLY).

``prove`` has the signature of ``oracle.prover.prove`` / ``luminair_b200.prover.prove`` and returns the bincode
``LuminairProof``.  It is the packed (AVX-512 / AVX2) + OpenMP restatement of the reference's CPU prover
(/root/reference/crates/prover/src/prover.rs:28-319 over stwo's SimdBackend with the "parallel" feature,
/root/reference/Cargo.toml:21-26): the checker for proofs at sizes the numpy oracle cannot reach, and the CPU baseline of
``bench.py``.  Pinned by ``tests/test_cpu_prover.py`` (the reference's committed proof, byte equality with the numpy oracle).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

SLOT_OF = {"add": 0, "mul": 1, "recip": 2, "sin": 3, "sin_lookup": 4, "sum_reduce": 5, "max_reduce": 6, "sqrt": 7, "rem": 8,
           "exp2": 9, "exp2_lookup": 10, "log2": 11, "log2_lookup": 12, "less_than": 13, "range_check_lookup": 14,
           "inputs": 15, "contiguous": 16}  # LuminairClaim field order, crates/air/src/lib.rs:30-48
_LUT_OF_PREFIX = {"sin_lut_": 1, "exp2_lut_": 2, "log2_lut_": 3, "range_check_": 4}
STAGE_NAMES = ["preprocessed + main trace: interpolate+LDE+Merkle", "interaction trace: LogUp+interpolate+LDE+Merkle",
               "constraint quotients + composition commit", "OODS sampling (eval_at_point)", "DEEP quotients",
               "FRI commit (folds + layer Merkle)", "grind + queries + decommitment", "OODS check"]


class _Table(C.Structure):
    _fields_ = [("slot", C.c_int), ("n_cols", C.c_int), ("n_rows", C.c_uint64), ("rows", C.c_void_p)]


class _Lut(C.Structure):
    _fields_ = [("lut", C.c_int), ("col_index", C.c_int), ("log_size", C.c_int), ("values", C.c_void_p)]


class _Config(C.Structure):
    _fields_ = [("pow_bits", C.c_uint32), ("log_blowup_factor", C.c_uint32), ("log_last_layer_degree_bound", C.c_uint32),
                ("n_queries", C.c_uint64), ("channel_variant", C.c_int), ("n_slots", C.c_int), ("air_era", C.c_int),
                ("draw_lookup_elements", C.c_int)]


class CpuProvingError(Exception):
    pass


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)


def lib():
    """The widest build this host can run: 16 lanes (AVX-512) or 8 (AVX2)."""
    global _lib
    if _lib is None:
        flags = _cpu_flags()
        v4 = {"avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512cd"} <= flags
        path = os.path.join(_HERE, "_c", "libcpu_prover_v4.so" if v4 else "libcpu_prover_v3.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.ocp_last_error.restype = C.c_char_p
        _lib.ocp_lanes.restype = C.c_int
        _lib.ocp_max_threads.restype = C.c_int
    return _lib


def host_cores() -> int:
    """Cores this process may use (ignores OMP_NUM_THREADS, which torchrun sets to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(n: int | None = None) -> int:
    n = n or host_cores()
    lib().ocp_set_threads(int(n))
    return n


def lanes() -> int:
    return lib().ocp_lanes()


def prove(pie, n_slots: int = 17, slot_of=SLOT_OF, channel_variant="legacy", config=None, draw_lookup_elements=True,
          preprocessed=(), air_era="current", n_threads: int | None = None, return_stages=False):
    """pie: [(name, rows[n, n_cols])]; preprocessed: [(id, values[2^k])] in ``lookups_to_preprocessed_column`` order.
    -> bincode bytes (and the per-stage wall-clock ms when return_stages)."""
    L = lib()
    set_threads(n_threads)
    keep = []
    tables = (_Table * max(len(pie), 1))()
    for i, (name, rows) in enumerate(pie):
        arr = np.ascontiguousarray(np.asarray(rows), dtype=np.uint32)
        if arr.ndim != 2:
            raise ValueError("trace table must be [n_rows, n_cols]")
        keep.append(arr)
        tables[i].slot = slot_of[name]
        tables[i].n_rows, tables[i].n_cols = arr.shape
        tables[i].rows = arr.ctypes.data
    luts = (_Lut * max(len(preprocessed), 1))()
    for i, (cid, values) in enumerate(preprocessed):
        arr = np.ascontiguousarray(np.asarray(values), dtype=np.uint32).reshape(-1)
        keep.append(arr)
        for prefix, lut in _LUT_OF_PREFIX.items():
            if cid.startswith(prefix):
                luts[i].lut, luts[i].col_index = lut, int(cid.rsplit("_", 1)[1])
                break
        else:
            raise ValueError(f"unknown preprocessed column id '{cid}'")
        luts[i].log_size = arr.size.bit_length() - 1
        luts[i].values = arr.ctypes.data
    pow_bits, blow, last, nq = (5, 1, 0, 3) if config is None else (
        config.pow_bits, config.log_blowup_factor, config.log_last_layer_degree_bound, config.n_queries)
    cfg = _Config(pow_bits, blow, last, nq, {"legacy": 0, "v2": 1}[channel_variant], n_slots,
                  {"current": 0, "artifact": 1}[air_era], 1 if draw_lookup_elements else 0)
    out = C.c_void_p()
    out_len = C.c_size_t()
    stages = (C.c_double * 8)()
    rc = L.ocp_prove(tables, len(pie), luts, len(preprocessed), C.byref(cfg), C.byref(out), C.byref(out_len), stages)
    if rc == -5:
        raise CpuProvingError(L.ocp_last_error().decode())
    if rc == -3:
        raise ValueError(L.ocp_last_error().decode())
    if rc != 0:
        raise RuntimeError(L.ocp_last_error().decode())
    try:
        data = C.string_at(out.value, out_len.value)
    finally:
        L.ocp_free(out)
    if return_stages:
        return data, dict(zip(STAGE_NAMES, [float(x) for x in stages]))
    return data


def cfft(values: np.ndarray, forward: bool, n_threads: int | None = None):
    """In-place transform of values[n_cols, 2^log] (uint32, C-contiguous): forward = coefficients -> evaluations on
    CanonicCoset(log).circle_domain(); else interpolation."""
    assert values.dtype == np.uint32 and values.flags.c_contiguous and values.ndim == 2
    set_threads(n_threads)
    log = values.shape[1].bit_length() - 1
    lib().ocp_cfft(values.ctypes.data_as(C.c_void_p), C.c_size_t(values.shape[1]), values.shape[0], log, 1 if forward else 0)


def merkle_root(columns) -> bytes:
    cols = [np.ascontiguousarray(c, dtype=np.uint32) for c in columns]
    ptrs = (C.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
    logs = (C.c_int * max(len(cols), 1))(*[c.size.bit_length() - 1 for c in cols])
    out = C.create_string_buffer(32)
    lib().ocp_merkle_root(ptrs, logs, len(cols), out)
    return out.raw


def eval_at_point(columns, point) -> np.ndarray:
    """columns: list of coefficient arrays of one size; point: 8 u32 (x then y).  -> [n_cols, 4] uint32."""
    cols = [np.ascontiguousarray(c, dtype=np.uint32) for c in columns]
    ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    pt = (C.c_uint32 * 8)(*[int(x) for x in point])
    out = np.empty((len(cols), 4), dtype=np.uint32)
    lib().ocp_eval_at_point(ptrs, len(cols), cols[0].size.bit_length() - 1, pt, out.ctypes.data_as(C.c_void_p))
    return out
