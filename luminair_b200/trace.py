"""gen_trace on the device: the operators' ``process_trace`` loops
(crates/graph/src/op/prim.rs, driven by ``LuminairGraph::gen_trace``, crates/graph/src/graph.rs:161-604) as CUDA kernels
behind ``lb_trace_op`` (every operator) and ``lb_trace_inputs / lb_trace_add / lb_trace_mul`` (the element-wise fast path).
Tensors stay in HBM as raw Fixed<12> int32 values; the trace tables are written there too (row-major, the format
``prove(device_tables=...)`` takes), so a proof of a graph moves only its input tensors over PCIe instead of the trace
tables (15x more bytes for a + b).

The graph is recorded first and emitted in ``finish()``, because the multiplicity a node yields an element with is the number
of rows that consume it (``node_info.num_consumers``, op/prim.rs:947-951), known only once the graph is complete.  Same
interface, node numbering, row order and table order as the host builder used by the tests (a numpy ``GraphTrace`` kept with the test infrastructure) (creation order;
tables in claim-slot order), which the tests compare against: operands are ``node`` or ``(node, gather index)``.

Lookup operators (sin / exp2 / log2) need the ``LookupLayout`` of the circuit settings (crates/air/src/preprocessed.rs:41-116,
produced by the reference's calibration pass ``gen_circuit_settings``): the LUT columns are generated on the host exactly as
the reference does (``lookups.lut_columns``: f64 libm + Fixed::from_f64), uploaded once, and the device reads ``f(x)`` from them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Lookup, LuminairB200Error, TraceOpDesc, check
from .backend import CudaBackend
from .lookups import RANGE_CHECK_BITS, lut_columns, range_check_column

P = (1 << 31) - 1
N_COLS = {"add": 15, "mul": 16, "recip": 13, "sin": 12, "sum_reduce": 14, "max_reduce": 15, "sqrt": 13, "rem": 16,
          "exp2": 12, "log2": 12, "less_than": 22, "inputs": 7, "contiguous": 11}
# LB_OP_* (include/luminair_b200.h): the claim slot of the operator's component
OP_CODE = {"add": 0, "mul": 1, "recip": 2, "sin": 3, "sum_reduce": 5, "max_reduce": 6, "sqrt": 7, "rem": 8, "exp2": 9,
           "log2": 11, "less_than": 13, "inputs": 15, "contiguous": 16}
ORDER = ["add", "mul", "recip", "sin", "sin_lookup", "sum_reduce", "max_reduce", "sqrt", "rem", "exp2", "exp2_lookup", "log2",
         "log2_lookup", "less_than", "range_check_lookup", "inputs", "contiguous"]
_LUT_HOST = {}  # (lut name, layout ranges) -> [(column id, values)]: lookups.lut_columns of a circuit-settings layout
LUT_OPS = ("sin", "exp2", "log2")


class DeviceGraphTrace:
    def __init__(self, backend: CudaBackend):
        self.be = backend
        self.nodes = []   # (kind, payload): ("inputs", device buffer) | (op, [(operand node, device idx buffer | None), ...], group)
        self.sizes = []   # node -> elements
        self.reads = []   # node -> [(device idx buffer | None, n_reads)] of its consumers
        self._keep = []
        self._program = None  # after the first finish(): the emission calls, replayed by later finish() calls

    # -- nodes ------------------------------------------------------------------------------------------------
    def input(self, raw_values) -> int:
        """Upload one input tensor (raw Fixed<12> values) - the only host->device traffic of the graph."""
        arr = np.asarray(raw_values)
        if arr.dtype != np.int32 or not arr.flags.c_contiguous:  # int32 arrays (e.g. pinned) are uploaded as they are
            arr = np.ascontiguousarray(arr.astype(np.int64).astype(np.int32))
        arr = arr.reshape(-1)
        buf = self.be.upload(arr.view(np.uint32), self.be.alloc(arr.size, pooled=True))
        return self._node("inputs", buf, arr.size)

    def set_input(self, node: int, raw_values):
        """New values for an input tensor of a recorded graph (same size); the next ``finish()`` re-runs the graph with them -
        the reference's compile-once / ``gen_trace``-per-execution split (StwoCompiler, then graph.rs:161-604 per run)."""
        kind, buf = self.nodes[node]
        arr = np.ascontiguousarray(np.asarray(raw_values).astype(np.int64).astype(np.int32)).reshape(-1)
        if kind != "inputs" or arr.size != self.sizes[node]:
            raise LuminairB200Error("set_input: not an input node, or a different number of elements")
        self.be.upload(arr.view(np.uint32), buf)

    def input_device(self, buf, n: int) -> int:
        """An input tensor already resident in HBM (int32 raw values)."""
        return self._node("inputs", buf, n)

    def _node(self, kind, payload, n):
        self.nodes.append((kind, payload))
        self.sizes.append(int(n))
        self.reads.append([])
        return len(self.nodes) - 1

    def _read(self, operand):
        """-> (node, device gather index or None, rows that read)."""
        node, idx = operand if isinstance(operand, tuple) else (operand, None)
        if idx is None:
            d_idx, n = None, self.sizes[node]
        else:
            idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int64).reshape(-1))
            if idx.size and (idx.min() < 0 or idx.max() >= self.sizes[node]):
                raise LuminairB200Error("gather index out of range")
            d_idx = self.be.upload(idx.astype(np.uint32), self.be.alloc(max(idx.size, 1), pooled=True))
            n = idx.size
        self.reads[node].append((d_idx, n))
        return node, d_idx, n

    def _op(self, kind, operands, group=1):
        srcs = [self._read(o) for o in operands]
        rows = srcs[0][2]
        if any(s[2] != rows for s in srcs):
            raise LuminairB200Error("element-wise operands must have the same number of elements")
        if rows % group:
            raise LuminairB200Error("reduction input is not a whole number of groups")
        return self._node(kind, ([(s[0], s[1]) for s in srcs], group), rows // group)

    def add(self, a, b): return self._op("add", [a, b])
    def mul(self, a, b): return self._op("mul", [a, b])
    def rem(self, a, b): return self._op("rem", [a, b])
    def less_than(self, a, b): return self._op("less_than", [a, b])
    def recip(self, a): return self._op("recip", [a])
    def sqrt(self, a): return self._op("sqrt", [a])
    def sin(self, a): return self._op("sin", [a])
    def exp2(self, a): return self._op("exp2", [a])
    def log2(self, a): return self._op("log2", [a])
    def contiguous(self, a): return self._op("contiguous", [a])
    def sum_reduce(self, a, group): return self._op("sum_reduce", [a], int(group))
    def max_reduce(self, a, group): return self._op("max_reduce", [a], int(group))

    # -- emission ---------------------------------------------------------------------------------------------
    def _zeros(self, n):
        buf = self.be.alloc(n, pooled=True)
        check(self.be.ctx, self.be.lib.lb_memset_zero(self.be.ctx, C.c_void_p(buf.ptr), n), "lb_memset_zero")
        return buf

    def finish(self, layouts=None):
        """Run the graph on the device.  ``layouts``: {"sin" | "exp2" | "log2": LookupLayout} of the circuit settings.
        -> (pie_meta, device_tables, values): ``pie_meta`` = [(name, None)] in claim-slot order for ``prove``,
        ``device_tables`` = {name: (ptr, n_rows, n_cols)}, ``values`` = device buffers of every node's tensor (int32 raw
        values).  ``self.preprocessed`` holds the LUT columns for ``prove(preprocessed=...)``."""
        be, lib, ctx = self.be, self.be.lib, self.be.ctx
        if self._program is not None:
            # the graph was planned by an earlier call (buffers, consumer counts, gather indices, descriptors): clear the lookup
            # counters and replay the emission calls on the current input values
            for _, _, d_mult, n_entries in self.lookups.values():
                check(ctx, lib.lb_memset_zero(ctx, C.c_void_p(d_mult.ptr), n_entries), "lb_memset_zero")
            for call in self._program:
                call()
            return self._result
        program = []

        def emit(fn, args, what):
            def call():
                check(ctx, fn(ctx, *args), what)
            program.append(call)
            call()

        layouts = layouts or {}
        rows_total = {}
        for (kind, payload), n in zip(self.nodes, self.sizes):
            rows_total[kind] = rows_total.get(kind, 0) + n * (payload[1] if kind != "inputs" else 1)
        tables = {k: be.alloc(rows_total[k] * N_COLS[k], pooled=True) for k in rows_total if rows_total[k]}

        # lookup tables: LUT columns from the host (as the reference generates them), multiplicity counters on the device
        self.preprocessed = []
        lookups = {}
        for name in LUT_OPS:
            if name not in rows_total:
                continue
            if name not in layouts:
                raise LuminairB200Error(f"'{name}' needs its LookupLayout (circuit settings)")
            layout = layouts[name]
            if len(layout.ranges) > 8:
                raise LuminairB200Error("at most 8 lookup ranges (LB_MAX_LOOKUP_RANGES)")
            # the columns depend only on the circuit settings: generated once per layout, the device copy once per backend
            key = (name, tuple(layout.ranges))
            cols = _LUT_HOST.get(key)
            if cols is None:
                cols = _LUT_HOST[key] = lut_columns(name, layout)
            self.preprocessed += cols
            lk = Lookup()
            lk.n_ranges = len(layout.ranges)
            for k, (lo, hi) in enumerate(layout.ranges):
                lk.lo[k], lk.hi[k] = lo, hi
            dev_cache = be.__dict__.setdefault("_lut_device_columns", {})
            d_vals = dev_cache.get(key)
            if d_vals is None:
                d_vals = dev_cache[key] = be.upload(cols[1][1])
            d_mult = self._zeros(1 << layout.log_size)
            lk.d_values, lk.d_multiplicities = d_vals.ptr, d_mult.ptr
            lookups[name] = (lk, d_vals, d_mult, 1 << layout.log_size)
        if "less_than" in rows_total:
            self.preprocessed += range_check_column()
            n_rc = 1 << max(RANGE_CHECK_BITS, 4)
            d_mult = self._zeros(n_rc)
            lk = Lookup()
            lk.d_multiplicities = d_mult.ptr
            lookups["less_than"] = (lk, None, d_mult, n_rc)

        at = {k: 0 for k in rows_total}
        values = []
        for node, ((kind, payload), n) in enumerate(zip(self.nodes, self.sizes)):
            reads = self.reads[node]
            identity_consumers = all(d_idx is None for d_idx, _ in reads)
            fast = kind in ("inputs", "add", "mul") and identity_consumers and (
                kind == "inputs" or all(d_idx is None for _, d_idx in payload[0]))
            if fast:  # scalar multiplicity = number of consumers (0 for a final output, op/prim.rs:947-951)
                mult = len(reads) % P
                if kind == "inputs":
                    buf = payload
                    emit(lib.lb_trace_inputs, (node, C.c_void_p(buf.ptr), n, mult, C.c_void_p(tables[kind].ptr), at[kind]),
                         "lb_trace_inputs")
                else:
                    (a, _), (b, _) = payload[0]
                    buf = be.alloc(max(n, 1), pooled=True)
                    fn = lib.lb_trace_add if kind == "add" else lib.lb_trace_mul
                    emit(fn, (node, a, b, C.c_void_p(values[a].ptr), C.c_void_p(values[b].ptr), n, mult, C.c_void_p(buf.ptr),
                              C.c_void_p(tables[kind].ptr), at[kind]), "lb_trace_" + kind)
                values.append(buf)
                at[kind] += n
                continue
            d_uses = None
            if reads:
                d_uses = self._zeros(max(n, 1))
                for d_idx, n_reads in reads:
                    check(ctx, lib.lb_trace_count_uses(ctx, C.c_void_p(d_uses.ptr), C.c_void_p(d_idx.ptr if d_idx else None),
                                                       n_reads), "lb_trace_count_uses")
                self._keep.append(d_uses)
            d = TraceOpDesc()
            d.op, d.node_id, d.n = OP_CODE[kind], node, n
            d.d_out_mult = d_uses.ptr if d_uses else None
            d.d_rows, d.row0 = tables[kind].ptr, at[kind]
            if kind == "inputs":
                buf = payload
                d.d_lhs, d.group = buf.ptr, 1
            else:
                srcs, group = payload
                buf = be.alloc(max(n, 1), pooled=True)
                d.d_out, d.group = buf.ptr, group
                d.lhs_id, d.d_lhs = srcs[0][0], values[srcs[0][0]].ptr
                d.d_lhs_idx = srcs[0][1].ptr if srcs[0][1] else None
                if len(srcs) > 1:
                    d.rhs_id, d.d_rhs = srcs[1][0], values[srcs[1][0]].ptr
                    d.d_rhs_idx = srcs[1][1].ptr if srcs[1][1] else None
                if kind in lookups:
                    d.lookup = C.pointer(lookups[kind][0])
            self._keep.append(d)
            emit(lib.lb_trace_op, (C.byref(d),), "lb_trace_op(" + kind + ")")
            values.append(buf)
            at[kind] += n * (payload[1] if kind != "inputs" else 1)

        device_tables = {k: (tables[k].ptr, rows_total[k], N_COLS[k]) for k in tables}
        for name, (_, _, d_mult, n_entries) in lookups.items():
            tname = "range_check_lookup" if name == "less_than" else name + "_lookup"
            tables[tname] = d_mult
            device_tables[tname] = (d_mult.ptr, n_entries, 1)
        order = [k for k in ORDER if k in device_tables]
        self.tables, self.values, self.lookups = tables, values, lookups  # the buffers must outlive the prove() call
        pie_meta = [(k, None) for k in order]  # table order for prove(); the rows live on the device (device_tables)
        self._program = program
        self._result = (pie_meta, {k: device_tables[k] for k in order}, values)
        return self._result
