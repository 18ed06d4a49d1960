"""ctypes binding of include/luminair_b200.h (no torch types cross this boundary)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LUMINAIR_B200_LIB") or os.path.join(_HERE, "libluminair_b200.so")

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
    "lb_alloc_pooled": (C.c_int, [ctxp, C.c_size_t, C.POINTER(C.c_void_p)]),
    "lb_free_pooled": (C.c_int, [ctxp, C.c_void_p]),
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


class TraceTable(C.Structure):
    """lb_trace_table"""
    _fields_ = [("slot", C.c_int), ("n_cols", C.c_int), ("n_rows", C.c_uint64), ("rows", C.c_void_p),
                ("rows_on_device", C.c_int)]


class ProveConfig(C.Structure):
    """lb_prove_config"""
    _fields_ = [("pow_bits", C.c_uint32), ("log_blowup_factor", C.c_uint32), ("log_last_layer_degree_bound", C.c_uint32),
                ("n_queries", C.c_uint64), ("channel_variant", C.c_int), ("n_slots", C.c_int), ("air_era", C.c_int),
                ("draw_lookup_elements", C.c_int)]


class PreprocessedColumn(C.Structure):
    """lb_preprocessed_column"""
    _fields_ = [("lut", C.c_int), ("col_index", C.c_int), ("log_size", C.c_int), ("values", C.c_void_p),
                ("on_device", C.c_int)]


class Lookup(C.Structure):
    """lb_lookup"""
    _fields_ = [("n_ranges", C.c_int), ("lo", C.c_int32 * 8), ("hi", C.c_int32 * 8), ("d_values", C.c_void_p),
                ("d_multiplicities", C.c_void_p)]


class TraceOpDesc(C.Structure):
    """lb_trace_op_desc"""
    _fields_ = [("op", C.c_int), ("node_id", C.c_uint32), ("lhs_id", C.c_uint32), ("rhs_id", C.c_uint32),
                ("d_lhs", C.c_void_p), ("d_lhs_idx", C.c_void_p), ("d_rhs", C.c_void_p), ("d_rhs_idx", C.c_void_p),
                ("n", C.c_uint64), ("group", C.c_uint32), ("d_out_mult", C.c_void_p), ("d_out", C.c_void_p),
                ("d_rows", C.c_void_p), ("row0", C.c_uint64), ("lookup", C.POINTER(Lookup))]


class Relation(C.Structure):
    """lb_relation"""
    _fields_ = [("z", C.c_uint32 * 4), ("alpha", C.c_uint32 * 4)]


class BatchShard(C.Structure):
    """lb_batch_shard"""
    _fields_ = [("col_offset", C.c_int), ("n_cols_global", C.c_int)]


class SampleBatch(C.Structure):
    """lb_sample_batch"""
    _fields_ = [("point", C.c_uint32 * 8), ("n_cols", C.c_int), ("col_idx", C.POINTER(C.c_int)),
                ("values", C.POINTER(C.c_uint32))]


SIGNATURES.update({
    "lb_eval_at_point": (C.c_int, [ctxp, C.POINTER(C.c_void_p), C.c_int, C.c_int, u32p, u32p]),
    "lb_accumulate_quotients": (C.c_int, [ctxp, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(SampleBatch), C.c_int,
                                          u32p, C.POINTER(C.c_void_p)]),
    "lb_accumulate_quotients_shard": (C.c_int, [ctxp, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(SampleBatch),
                                                C.POINTER(BatchShard), C.c_int, u32p, C.POINTER(C.c_void_p)]),
    "lb_fold_circle_into_line": (C.c_int, [ctxp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, u32p]),
    "lb_fold_line": (C.c_int, [ctxp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, u32p]),
    "lb_grind": (C.c_int, [ctxp, u32p, C.c_int, C.c_uint32, C.POINTER(C.c_uint64)]),
    "lb_logup_interaction_trace": (C.c_int, [ctxp, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, u32p,
                                             u32p, u32p]),
    "lb_constraint_quotients": (C.c_int, [ctxp, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, u32p, u32p,
                                          u32p, u32p, C.c_int, C.POINTER(C.c_void_p), C.c_int]),
    "lb_logup_interaction_trace_lut": (C.c_int, [ctxp, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p,
                                                 C.c_size_t, C.c_int, C.POINTER(Relation), u32p]),
    "lb_constraint_quotients_lut": (C.c_int, [ctxp, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                              C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(Relation), u32p, u32p, C.c_int,
                                              C.POINTER(C.c_void_p), C.c_int]),
    "lb_prove_with_lookups": (C.c_int, [ctxp, C.POINTER(TraceTable), C.c_int, C.POINTER(PreprocessedColumn), C.c_int,
                                        C.POINTER(ProveConfig), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "lb_evaluate_batch_scatter": (C.c_int, [ctxp, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                            C.POINTER(C.c_void_p), C.c_int, C.c_size_t]),
    "lb_ipc_export": (C.c_int, [ctxp, C.c_void_p, C.c_char_p]),
    "lb_ipc_open": (C.c_int, [ctxp, C.c_char_p, C.POINTER(C.c_void_p)]),
    "lb_ipc_close": (C.c_int, [ctxp, C.c_void_p]),
    "lb_lde_host": (C.c_int, [ctxp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "lb_prove": (C.c_int, [ctxp, C.POINTER(TraceTable), C.c_int, C.POINTER(ProveConfig), C.POINTER(C.c_void_p),
                           C.POINTER(C.c_size_t)]),
    "lb_free_host": (None, [C.c_void_p]),
    "lb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "lb_comm_init": (C.c_int, [ctxp, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "lb_comm_destroy": (None, [C.c_void_p]),
    "lb_comm_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                C.POINTER(C.c_int)]),
    "lb_prove_sharded": (C.c_int, [ctxp, C.c_void_p, C.POINTER(TraceTable), C.c_int, C.POINTER(PreprocessedColumn), C.c_int,
                                   C.POINTER(ProveConfig), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "lb_trace_inputs": (C.c_int, [ctxp, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64]),
    "lb_trace_add": (C.c_int, [ctxp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                               C.c_void_p, C.c_void_p, C.c_uint64]),
    "lb_trace_mul": (C.c_int, [ctxp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                               C.c_void_p, C.c_void_p, C.c_uint64]),
    "lb_bit_reverse": (C.c_int, [ctxp, C.c_void_p, C.c_int]),
    "lb_new_canonical_ordered": (C.c_int, [ctxp, C.c_void_p, C.c_void_p, C.c_int]),
    "lb_batch_inverse_m31": (C.c_int, [ctxp, C.c_void_p, C.c_void_p, C.c_size_t]),
    "lb_batch_inverse_qm31": (C.c_int, [ctxp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t]),
    "lb_accumulate": (C.c_int, [ctxp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t]),
    "lb_generate_secure_powers": (C.c_int, [u32p, C.c_int, u32p]),
    "lb_trace_op": (C.c_int, [ctxp, C.POINTER(TraceOpDesc)]),
    "lb_trace_count_uses": (C.c_int, [ctxp, C.c_void_p, C.c_void_p, C.c_uint64]),
    "lb_prove_transcript": (C.c_int, [ctxp, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "lb_prove_stage_ms": (C.c_int, [ctxp, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
})


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
