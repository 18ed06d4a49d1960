"""prove(): host mirror of ``luminair_prover::prover::prove``
(crates/prover/src/prover.rs:28-319) over the C ABI entry ``lb_prove``.

The pie is a list of trace tables in ``LuminairPie.trace_tables`` order
(crates/air/src/pie.rs:143-148): ``(name, rows)`` with ``rows`` an ``[n_rows, n_cols]`` array of
canonical M31 values, one row per ``*TraceTableRow`` (e.g. add/table.rs).  The result is the
bincode ``LuminairProof`` (crates/prover/src/lib.rs:22-32).  Everything runs on the GPU; there
is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import LuminairB200Error, PreprocessedColumn, ProveConfig, TraceTable, check
from .backend import CudaBackend

# field index of each component in LuminairClaim (crates/air/src/lib.rs:30-48)
CLAIM_SLOT = {"add": 0, "mul": 1, "recip": 2, "sin": 3, "sin_lookup": 4, "sum_reduce": 5, "max_reduce": 6, "sqrt": 7,
              "rem": 8, "exp2": 9, "exp2_lookup": 10, "log2": 11, "log2_lookup": 12, "less_than": 13,
              "range_check_lookup": 14, "inputs": 15, "contiguous": 16}
N_CLAIM_SLOTS = 17
# LUT of a preprocessed column id (crates/air/src/preprocessed.rs:289-296, 336-340 and siblings) -> LB_REL_*
LUT_OF_PREFIX = {"sin_lut_": 1, "exp2_lut_": 2, "log2_lut_": 3, "range_check_": 4}


def _lut_column(cid: str):
    for prefix, lut in LUT_OF_PREFIX.items():
        if cid.startswith(prefix):
            return lut, int(cid.rsplit("_", 1)[1])
    raise LuminairB200Error(f"unknown preprocessed column id '{cid}'")


class ProvingError(LuminairB200Error):
    """LuminairError::ProverError (crates/utils/src/lib.rs:17-18)."""


class TraceError(LuminairB200Error):
    """TraceError::EmptyTrace (e.g. add/witness.rs:39-41)."""


@dataclass
class PcsConfig:
    """stwo PcsConfig::default() as used at prover.rs:36."""
    pow_bits: int = 5
    log_blowup_factor: int = 1
    log_last_layer_degree_bound: int = 0
    n_queries: int = 3


class Comm:
    """``lb_comm``: this rank's handle of the multi-GPU prover (one process per GPU, NCCL over NVLink).  ``unique_id()`` is
    created on rank 0 and handed to the other ranks by the host's own means (``torch.distributed`` broadcast, MPI, a file)."""

    def __init__(self, backend: CudaBackend, unique_id: bytes, rank: int, world: int):
        self.be = backend
        h = C.c_void_p()
        check(backend.ctx, backend.lib.lb_comm_init(backend.ctx, unique_id, rank, world, C.byref(h)), "lb_comm_init")
        self.handle, self.rank, self.world = h, rank, world

    @staticmethod
    def unique_id(backend: CudaBackend) -> bytes:
        buf = C.create_string_buffer(128)
        rc = backend.lib.lb_comm_unique_id(buf)
        if rc != 0:
            raise LuminairB200Error(f"lb_comm_unique_id failed ({rc}): NCCL not available")
        return buf.raw

    def stats(self):
        sent, recv, n = C.c_uint64(), C.c_uint64(), C.c_int()
        self.be.lib.lb_comm_stats(self.handle, None, None, C.byref(sent), C.byref(recv), C.byref(n))
        return {"bytes_sent": sent.value, "bytes_received": recv.value, "n_collectives": n.value}

    def close(self):
        if self.handle:
            self.be.lib.lb_comm_destroy(self.handle)
            self.handle = None


def prove(pie, backend: CudaBackend | None = None, config: PcsConfig | None = None, channel_variant: str = "legacy",
          claim_slots=CLAIM_SLOT, n_slots: int = N_CLAIM_SLOTS, device_tables=None, air_era: str = "current",
          preprocessed=(), settings=None, comm: Comm | None = None) -> bytes:
    """pie: [(name, rows[n_rows, n_cols])].  device_tables: optional {name: (device_ptr, n_rows, n_cols)} to prove
    from tables already resident in HBM (bench.py's device-resident leg).  preprocessed: [(id, values[2^k])] LUT
    columns of the circuit settings in ``lookups_to_preprocessed_column`` order (preprocessed.rs:181-206); or pass
    ``settings`` (``luminair_b200.settings.CircuitSettings``), as the reference's ``prove(pie, settings)`` does
    (prover.rs:28-31), and the columns are derived from its lookup layouts.  comm: a ``Comm`` - the proof is then made by
    all its ranks together (``lb_prove_sharded``; every rank passes the same arguments and receives the same bytes)."""
    if settings is not None:
        if preprocessed:
            raise LuminairB200Error("pass either `settings` or `preprocessed`, not both")
        preprocessed = settings.preprocessed_columns()
    own = backend is None
    be = backend or CudaBackend(0)
    try:
        config = config or PcsConfig()
        n = len(pie)
        tables = (TraceTable * max(n, 1))()
        keep = []
        for i, (name, rows) in enumerate(pie):
            if name not in claim_slots:
                raise LuminairB200Error(f"component '{name}' is not supported by the B200 backend yet")
            tables[i].slot = claim_slots[name]
            if device_tables and name in device_tables:
                ptr, n_rows, n_cols = device_tables[name]
                tables[i].n_rows, tables[i].n_cols = n_rows, n_cols
                tables[i].rows = ptr
                tables[i].rows_on_device = 1
                continue
            arr = np.ascontiguousarray(np.asarray(rows), dtype=np.uint32)
            if arr.ndim != 2:
                raise LuminairB200Error("trace table must be [n_rows, n_cols]")
            if arr.shape[0] == 0:
                raise TraceError("EmptyTrace")
            keep.append(arr)
            tables[i].n_rows, tables[i].n_cols = arr.shape
            tables[i].rows = arr.ctypes.data
            tables[i].rows_on_device = 0
        cfg = ProveConfig(config.pow_bits, config.log_blowup_factor, config.log_last_layer_degree_bound, config.n_queries,
                          {"legacy": 0, "v2": 1}[channel_variant], n_slots, {"current": 0, "artifact": 1}[air_era], 1)
        n_pre = len(preprocessed)
        pre = (PreprocessedColumn * max(n_pre, 1))()
        for i, (cid, values) in enumerate(preprocessed):
            arr = np.ascontiguousarray(np.asarray(values), dtype=np.uint32).reshape(-1)
            if arr.size < 16 or arr.size & (arr.size - 1):
                raise LuminairB200Error("LUT column length must be a power of two >= 16")
            keep.append(arr)
            pre[i].lut, pre[i].col_index = _lut_column(cid)
            pre[i].log_size = arr.size.bit_length() - 1
            pre[i].values = arr.ctypes.data
            pre[i].on_device = 0
        out = C.c_void_p()
        out_len = C.c_size_t()
        if comm is not None:  # every rank calls this with the same pie; every rank gets the same bytes (lb_prove_sharded)
            rc = be.lib.lb_prove_sharded(be.ctx, comm.handle, tables, n, pre, n_pre, C.byref(cfg), C.byref(out), C.byref(out_len))
        else:
            rc = be.lib.lb_prove_with_lookups(be.ctx, tables, n, pre, n_pre, C.byref(cfg), C.byref(out), C.byref(out_len))
        if rc == -5:
            raise ProvingError(be.lib.lb_last_error(be.ctx).decode())
        check(be.ctx, rc, "lb_prove")
        try:
            return C.string_at(out.value, out_len.value)
        finally:
            be.lib.lb_free_host(out)
    finally:
        if own:
            be.close()


def last_transcript(be: CudaBackend):
    """Channel digest after every mix of the last prove() on this backend (diagnostics)."""
    n = C.c_size_t()
    be.lib.lb_prove_transcript(be.ctx, None, 0, C.byref(n))
    buf = C.create_string_buffer(32 * max(n.value, 1))
    be.lib.lb_prove_transcript(be.ctx, buf, n.value, C.byref(n))
    return [buf.raw[32 * i: 32 * i + 32] for i in range(n.value)]


def last_stage_ms(be: CudaBackend):
    n = C.c_int()
    be.lib.lb_prove_stage_ms(be.ctx, None, 0, C.byref(n))
    arr = (C.c_float * max(n.value, 1))()
    be.lib.lb_prove_stage_ms(be.ctx, arr, n.value, C.byref(n))
    return [float(arr[i]) for i in range(n.value)]


STAGE_NAMES = ["main trace: upload+interpolate+LDE+Merkle", "interaction trace: LogUp+interpolate+LDE+Merkle",
               "constraint quotients + composition commit", "OODS sampling (eval_at_point)", "DEEP quotients",
               "FRI commit (folds + layer Merkle)", "grind + queries + decommitment gathers", "OODS check + bincode"]
