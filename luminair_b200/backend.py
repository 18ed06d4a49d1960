"""CudaBackend: host mirror of the stwo backend traits LuminAIR names as ``SimdBackend``
(crates/prover/src/prover.rs:22,38,312; crates/air/src/utils.rs:112-128).

Every method is a thin call into the C ABI (include/luminair_b200.h); arrays that cross
the boundary are numpy uint32 host buffers or ``DeviceBuffer`` handles.  There is no CPU
path: constructing a backend without the CUDA library / a GPU raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import LuminairB200Error, check, load_library


class DeviceBuffer:
    """Owned device allocation of ``n`` u32 (stwo ``Column<M31>`` storage)."""

    def __init__(self, backend: "CudaBackend", n: int, pooled: bool = False):
        self.backend = backend
        self.n = int(n)
        self.pooled = pooled  # stream-ordered allocation from the device pool (lb_alloc_pooled)
        p = C.c_void_p()
        fn = backend.lib.lb_alloc_pooled if pooled else backend.lib.lb_alloc
        check(backend.ctx, fn(backend.ctx, self.n, C.byref(p)), "lb_alloc")
        self.ptr = p.value or 0

    def free(self):
        if self.ptr and self.backend.ctx:
            fn = self.backend.lib.lb_free_pooled if self.pooled else self.backend.lib.lb_free
            fn(self.backend.ctx, C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def at(self, offset_u32: int) -> int:
        return self.ptr + 4 * int(offset_u32)


class ColumnBatch:
    """n_cols columns of 2^log_size u32 at a fixed stride inside one DeviceBuffer."""

    def __init__(self, buf: DeviceBuffer, n_cols: int, log_size: int, stride: int | None = None, offset: int = 0):
        self.buf, self.n_cols, self.log_size = buf, n_cols, log_size
        self.stride = stride if stride is not None else (1 << log_size)
        self.offset = offset

    @property
    def ptr(self):
        return self.buf.at(self.offset)

    def col_ptr(self, c: int) -> int:
        return self.buf.at(self.offset + c * self.stride)

    def col_ptrs(self):
        return [self.col_ptr(c) for c in range(self.n_cols)]


class CudaBackend:
    def __init__(self, device: int = 0):
        self.lib = load_library()
        ctx = C.c_void_p()
        rc = self.lib.lb_ctx_create(device, C.byref(ctx))
        if rc != 0:
            raise LuminairB200Error(f"lb_ctx_create(device={device}) failed ({rc}): no usable CUDA device; "
                                    "luminair_b200 has no CPU fallback")
        self.ctx = ctx
        sm = C.c_int()
        mem = C.c_size_t()
        self.lib.lb_device_info(self.ctx, C.byref(sm), C.byref(mem))
        self.sm_count, self.total_mem = sm.value, mem.value

    def close(self):
        if self.ctx:
            self.lib.lb_ctx_destroy(self.ctx)
            self.ctx = None

    # ---- memory -------------------------------------------------------------------
    def alloc(self, n: int, pooled: bool = False) -> DeviceBuffer:
        return DeviceBuffer(self, n, pooled)

    def upload(self, arr: np.ndarray, buf: DeviceBuffer | None = None, offset: int = 0) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr, dtype=np.uint32)
        if buf is None:
            buf = self.alloc(arr.size)
        check(self.ctx, self.lib.lb_upload(self.ctx, C.c_void_p(buf.at(offset)), arr.ctypes.data_as(C.c_void_p), arr.size), "lb_upload")
        return buf

    def download(self, buf: DeviceBuffer, n: int | None = None, offset: int = 0) -> np.ndarray:
        n = buf.n - offset if n is None else n
        out = np.empty(n, dtype=np.uint32)
        check(self.ctx, self.lib.lb_download(self.ctx, out.ctypes.data_as(C.c_void_p), C.c_void_p(buf.at(offset)), n), "lb_download")
        return out

    def download_ptr(self, ptr: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint32)
        check(self.ctx, self.lib.lb_download(self.ctx, out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), n), "lb_download")
        return out

    def sync(self):
        check(self.ctx, self.lib.lb_sync(self.ctx), "lb_sync")

    def timer_start(self):
        check(self.ctx, self.lib.lb_timer_start(self.ctx), "lb_timer_start")

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        check(self.ctx, self.lib.lb_timer_stop_ms(self.ctx, C.byref(ms)), "lb_timer_stop_ms")
        return float(ms.value)

    # ---- PolyOps --------------------------------------------------------------------
    def precompute_twiddles(self, max_log: int):
        check(self.ctx, self.lib.lb_twiddles_ensure(self.ctx, max_log), "lb_twiddles_ensure")

    def export_twiddles(self, root_log: int) -> np.ndarray:
        self.precompute_twiddles(root_log + 1)
        buf = self.alloc(1 << root_log)
        check(self.ctx, self.lib.lb_twiddles_export(self.ctx, root_log, C.c_void_p(buf.ptr)), "lb_twiddles_export")
        out = self.download(buf)
        buf.free()
        return out

    def interpolate(self, cols: ColumnBatch):
        check(self.ctx, self.lib.lb_interpolate_batch(self.ctx, C.c_void_p(cols.ptr), cols.stride, cols.n_cols, cols.log_size), "lb_interpolate_batch")

    def evaluate(self, coeffs: ColumnBatch, out: ColumnBatch):
        assert coeffs.n_cols == out.n_cols
        check(self.ctx, self.lib.lb_evaluate_batch(self.ctx, C.c_void_p(coeffs.ptr), coeffs.stride, coeffs.log_size,
                                                    C.c_void_p(out.ptr), out.stride, out.log_size, out.n_cols), "lb_evaluate_batch")

    def lde_host(self, values: np.ndarray, log_out: int | None = None, out: np.ndarray | None = None,
                 coeffs_out: np.ndarray | None = None, chunk_cols: int = 0) -> np.ndarray:
        """values [n_cols, 2^log] (host) -> evaluations [n_cols, 2^log_out] (host); chunked, copy/compute overlapped."""
        assert values.dtype == np.uint32 and values.ndim == 2 and values.flags.c_contiguous
        n_cols, n = values.shape
        log_in = n.bit_length() - 1
        log_out = log_in if log_out is None else log_out
        if out is None:
            out = np.empty((n_cols, 1 << log_out), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.shape == (n_cols, 1 << log_out) and out.flags.c_contiguous
        cptr = C.c_void_p(coeffs_out.ctypes.data) if coeffs_out is not None else C.c_void_p(0)
        check(self.ctx, self.lib.lb_lde_host(self.ctx, C.c_void_p(values.ctypes.data), C.c_void_p(out.ctypes.data), n_cols,
                                              log_in, log_out, cptr, chunk_cols), "lb_lde_host")
        return out

    # ---- MerkleOps -------------------------------------------------------------------
    def merkle_commit_layer(self, log_size: int, prev_ptr: int | None, col_ptrs, out_ptr: int):
        n = len(col_ptrs)
        arr = (C.c_void_p * max(n, 1))(*col_ptrs)
        check(self.ctx, self.lib.lb_merkle_commit_layer(self.ctx, log_size, C.c_void_p(prev_ptr or 0), arr, n, C.c_void_p(out_ptr)), "lb_merkle_commit_layer")

    def gather_rows(self, col_ptrs, idx) -> np.ndarray:
        n = len(col_ptrs)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        out = np.empty((len(idx), n), dtype=np.uint32)
        if n == 0 or len(idx) == 0:
            return out
        arr = (C.c_void_p * n)(*col_ptrs)
        check(self.ctx, self.lib.lb_gather_rows(self.ctx, arr, n, idx.ctypes.data_as(C.c_void_p), len(idx), out.ctypes.data_as(C.c_void_p)), "lb_gather_rows")
        return out

    # ---- PolyOps::eval_at_point ---------------------------------------------------------
    def eval_at_point(self, col_ptrs, log_size: int, point) -> np.ndarray:
        """point = 8 u32 (x then y, QM31 coordinates) -> [n_cols, 4] u32."""
        n = len(col_ptrs)
        arr = (C.c_void_p * max(n, 1))(*col_ptrs)
        pt = (C.c_uint32 * 8)(*[int(v) for v in point])
        out = np.zeros((n, 4), dtype=np.uint32)
        check(self.ctx, self.lib.lb_eval_at_point(self.ctx, arr, n, log_size, pt, out.ctypes.data_as(C.POINTER(C.c_uint32))),
              "lb_eval_at_point")
        return out

    # ---- QuotientOps ---------------------------------------------------------------------
    def accumulate_quotients(self, log_size: int, col_ptrs, batches, random_coeff, out_ptrs, shards=None):
        """batches: list of (point[8], [(col_idx, value[4])...]); out_ptrs: 4 device coordinate columns.
        shards: optional [(col_offset, n_cols_global)] per batch -> this device's partial sum (lb_accumulate_quotients_shard)."""
        from ._lib import BatchShard, SampleBatch
        n = len(col_ptrs)
        arr = (C.c_void_p * n)(*col_ptrs)
        keep = []
        sb = (SampleBatch * len(batches))()
        for i, (pt, cols) in enumerate(batches):
            idx = (C.c_int * len(cols))(*[int(c) for c, _ in cols])
            vals = (C.c_uint32 * (4 * len(cols)))(*[int(x) for _, v in cols for x in v])
            keep += [idx, vals]
            sb[i].point = (C.c_uint32 * 8)(*[int(v) for v in pt])
            sb[i].n_cols = len(cols)
            sb[i].col_idx = idx
            sb[i].values = vals
        rc = (C.c_uint32 * 4)(*[int(v) for v in random_coeff])
        outs = (C.c_void_p * 4)(*out_ptrs)
        if shards is not None:
            sh = (BatchShard * len(batches))()
            for i, (off, tot) in enumerate(shards):
                sh[i].col_offset, sh[i].n_cols_global = int(off), int(tot)
            check(self.ctx, self.lib.lb_accumulate_quotients_shard(self.ctx, log_size, arr, n, sb, sh, len(batches), rc, outs),
                  "lb_accumulate_quotients_shard")
            return
        check(self.ctx, self.lib.lb_accumulate_quotients(self.ctx, log_size, arr, n, sb, len(batches), rc, outs),
              "lb_accumulate_quotients")

    # ---- FriOps ----------------------------------------------------------------------------
    def fold_circle_into_line(self, dst_ptrs, src_ptrs, log_size: int, alpha):
        d = (C.c_void_p * 4)(*dst_ptrs)
        s = (C.c_void_p * 4)(*src_ptrs)
        a = (C.c_uint32 * 4)(*[int(v) for v in alpha])
        check(self.ctx, self.lib.lb_fold_circle_into_line(self.ctx, d, s, log_size, a), "lb_fold_circle_into_line")

    def fold_line(self, dst_ptrs, src_ptrs, log_size: int, alpha):
        d = (C.c_void_p * 4)(*dst_ptrs)
        s = (C.c_void_p * 4)(*src_ptrs)
        a = (C.c_uint32 * 4)(*[int(v) for v in alpha])
        check(self.ctx, self.lib.lb_fold_line(self.ctx, d, s, log_size, a), "lb_fold_line")

    # ---- GrindOps ---------------------------------------------------------------------------
    def grind(self, digest: bytes, pow_bits: int, channel_variant: int = 0) -> int:
        dg = (C.c_uint32 * 8).from_buffer_copy(digest)
        nonce = C.c_uint64()
        check(self.ctx, self.lib.lb_grind(self.ctx, dg, channel_variant, pow_bits, C.byref(nonce)), "lb_grind")
        return int(nonce.value)

    # ---- ColumnOps / FieldOps / AccumulationOps ---------------------------------------------------
    @staticmethod
    def _ptr4(ptrs):
        return (C.c_void_p * 4)(*[C.c_void_p(int(p)) for p in ptrs])

    def bit_reverse_column(self, ptr: int, log_size: int):
        check(self.ctx, self.lib.lb_bit_reverse(self.ctx, C.c_void_p(ptr), log_size), "lb_bit_reverse")

    def new_canonical_ordered(self, src_ptr: int, dst_ptr: int, log_size: int):
        check(self.ctx, self.lib.lb_new_canonical_ordered(self.ctx, C.c_void_p(src_ptr), C.c_void_p(dst_ptr), log_size),
              "lb_new_canonical_ordered")

    def batch_inverse(self, src_ptr: int, dst_ptr: int, n: int):
        check(self.ctx, self.lib.lb_batch_inverse_m31(self.ctx, C.c_void_p(src_ptr), C.c_void_p(dst_ptr), n), "lb_batch_inverse_m31")

    def batch_inverse_secure(self, src_ptrs, dst_ptrs, n: int):
        check(self.ctx, self.lib.lb_batch_inverse_qm31(self.ctx, self._ptr4(src_ptrs), self._ptr4(dst_ptrs), n),
              "lb_batch_inverse_qm31")

    def accumulate(self, column_ptrs, other_ptrs, n: int):
        check(self.ctx, self.lib.lb_accumulate(self.ctx, self._ptr4(column_ptrs), self._ptr4(other_ptrs), n), "lb_accumulate")

    def generate_secure_powers(self, felt, n_powers: int) -> np.ndarray:
        f = (C.c_uint32 * 4)(*[int(x) for x in felt])
        out = np.zeros((max(n_powers, 1), 4), dtype=np.uint32)
        rc = self.lib.lb_generate_secure_powers(f, n_powers, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        if rc != 0:
            raise LuminairB200Error(f"lb_generate_secure_powers: {rc}")
        return out[:n_powers]

    # ---- AIR kernels -------------------------------------------------------------------------
    def logup_interaction_trace(self, component: int, main: ColumnBatch, inter: ColumnBatch, z, alpha) -> np.ndarray:
        zz = (C.c_uint32 * 4)(*[int(v) for v in z])
        aa = (C.c_uint32 * 4)(*[int(v) for v in alpha])
        claimed = (C.c_uint32 * 4)()
        check(self.ctx, self.lib.lb_logup_interaction_trace(self.ctx, component, C.c_void_p(main.ptr), main.stride,
                                                             C.c_void_p(inter.ptr), inter.stride, main.log_size, zz, aa, claimed),
              "lb_logup_interaction_trace")
        return np.array(list(claimed), dtype=np.uint32)

    @staticmethod
    def _relations(rels):
        """rels: 5 (z, alpha) pairs of 4 u32 in LB_REL_* order -> lb_relation[5]"""
        from ._lib import Relation
        arr = (Relation * 5)()
        for k, (z, alpha) in enumerate(rels):
            arr[k].z = (C.c_uint32 * 4)(*[int(v) for v in z])
            arr[k].alpha = (C.c_uint32 * 4)(*[int(v) for v in alpha])
        return arr

    def logup_interaction_trace_lut(self, component: int, main: ColumnBatch, inter: ColumnBatch, rels, lut_ptrs=()) -> np.ndarray:
        """Any component: rels = all 5 relations; lut_ptrs = device pointers of the LUT columns a table component reads."""
        claimed = (C.c_uint32 * 4)()
        lut = (C.c_void_p * 2)(*(list(lut_ptrs) + [None] * (2 - len(lut_ptrs))))
        check(self.ctx, self.lib.lb_logup_interaction_trace_lut(self.ctx, component, C.c_void_p(main.ptr), main.stride, lut,
                                                                 C.c_void_p(inter.ptr), inter.stride, main.log_size,
                                                                 self._relations(rels), claimed),
              "lb_logup_interaction_trace_lut")
        return np.array(list(claimed), dtype=np.uint32)

    def constraint_quotients_lut(self, component: int, main_ev: ColumnBatch, inter_ev: ColumnBatch, log_size: int, eval_log: int,
                                 rels, claimed_sum, pows, acc_ptrs, lut_ptrs=(), accumulate: bool = False):
        cs = (C.c_uint32 * 4)(*[int(v) for v in claimed_sum])
        flat = [int(x) for p in pows for x in p]
        pw = (C.c_uint32 * len(flat))(*flat)
        acc = (C.c_void_p * 4)(*acc_ptrs)
        lut = (C.c_void_p * 2)(*(list(lut_ptrs) + [None] * (2 - len(lut_ptrs))))
        check(self.ctx, self.lib.lb_constraint_quotients_lut(self.ctx, component, C.c_void_p(main_ev.ptr), main_ev.stride,
                                                              C.c_void_p(inter_ev.ptr), inter_ev.stride, lut, log_size, eval_log,
                                                              self._relations(rels), cs, pw, len(pows), acc,
                                                              1 if accumulate else 0),
              "lb_constraint_quotients_lut")

    def constraint_quotients(self, component: int, main_lde: ColumnBatch, inter_lde: ColumnBatch, log_size: int, z, alpha,
                             claimed_sum, pows, acc_ptrs, accumulate: bool = False):
        zz = (C.c_uint32 * 4)(*[int(v) for v in z])
        aa = (C.c_uint32 * 4)(*[int(v) for v in alpha])
        cs = (C.c_uint32 * 4)(*[int(v) for v in claimed_sum])
        flat = [int(x) for p in pows for x in p]
        pw = (C.c_uint32 * len(flat))(*flat)
        acc = (C.c_void_p * 4)(*acc_ptrs)
        check(self.ctx, self.lib.lb_constraint_quotients(self.ctx, component, C.c_void_p(main_lde.ptr), main_lde.stride,
                                                          C.c_void_p(inter_lde.ptr), inter_lde.stride, log_size, zz, aa, cs, pw,
                                                          len(pows), acc, 1 if accumulate else 0),
              "lb_constraint_quotients")
