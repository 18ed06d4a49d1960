"""gen_trace on the device for element-wise graphs: the operators' ``process_trace`` loops
(crates/graph/src/op/prim.rs: CopyToStwo :72-84, LuminairAdd :919-1013, LuminairMul) as CUDA kernels behind
``lb_trace_inputs / lb_trace_add / lb_trace_mul``.  Tensors stay in HBM as raw Fixed<12> int32 values; the trace tables are
written there too (row-major, the format ``prove(device_tables=...)`` takes), so a proof of ``f(a, b)`` moves only ``a`` and
``b`` over PCIe instead of the trace tables (15x more bytes for a + b).

The graph is recorded first and emitted in ``finish()``, because a node's yield multiplicity is its number of consumers
(``node_info.num_consumers``, op/prim.rs:947-951), known only once the graph is complete.  Same node numbering, row order and
table order as ``luminair_b200.pie.GraphTrace`` (creation order; tables in claim-slot order), which the tests compare against.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import LuminairB200Error, check
from .backend import CudaBackend

P = (1 << 31) - 1
N_COLS = {"add": 15, "mul": 16, "inputs": 7}


class DeviceGraphTrace:
    def __init__(self, backend: CudaBackend):
        self.be = backend
        self.nodes = []  # (kind, payload): ("inputs", device buffer) | ("add"/"mul", (lhs, rhs))
        self.sizes = []
        self.consumers = []

    def input(self, raw_values) -> int:
        """Upload one input tensor (raw Fixed<12> values) - the only host->device traffic of the graph."""
        arr = np.asarray(raw_values)
        if arr.dtype != np.int32 or not arr.flags.c_contiguous:  # int32 arrays (e.g. pinned) are uploaded as they are
            arr = np.ascontiguousarray(arr.astype(np.int64).astype(np.int32))
        arr = arr.reshape(-1)
        buf = self.be.upload(arr.view(np.uint32), self.be.alloc(arr.size, pooled=True))
        return self._node("inputs", buf, arr.size)

    def input_device(self, buf, n: int) -> int:
        """An input tensor already resident in HBM (int32 raw values)."""
        return self._node("inputs", buf, n)

    def _node(self, kind, payload, n):
        self.nodes.append((kind, payload))
        self.sizes.append(int(n))
        self.consumers.append(0)
        return len(self.nodes) - 1

    def _binary(self, kind, a, b):
        if self.sizes[a] != self.sizes[b]:
            raise LuminairB200Error("element-wise operands must have the same number of elements")
        self.consumers[a] += 1
        self.consumers[b] += 1
        return self._node(kind, (a, b), self.sizes[a])

    def add(self, a: int, b: int) -> int:
        return self._binary("add", a, b)

    def mul(self, a: int, b: int) -> int:
        return self._binary("mul", a, b)

    def finish(self):
        """Run the graph on the device.  -> (pie_meta, device_tables, values): ``pie_meta`` = [(name, None)]
        in claim-slot order for ``prove``, ``device_tables`` = {name: (ptr, n_rows, n_cols)}, ``values`` = device buffers of
        every node's tensor (int32 raw values)."""
        be, lib, ctx = self.be, self.be.lib, self.be.ctx
        rows_total = {k: 0 for k in N_COLS}
        for (kind, _), n in zip(self.nodes, self.sizes):
            rows_total[kind] += n
        tables = {k: be.alloc(rows_total[k] * N_COLS[k], pooled=True) for k in N_COLS if rows_total[k]}
        at = {k: 0 for k in N_COLS}
        values = []
        for node, ((kind, payload), n) in enumerate(zip(self.nodes, self.sizes)):
            mult = self.consumers[node] % P  # 0 for a final output (op/prim.rs:947-951)
            if kind == "inputs":
                buf = payload
                check(ctx, lib.lb_trace_inputs(ctx, node, C.c_void_p(buf.ptr), n, mult, C.c_void_p(tables[kind].ptr), at[kind]),
                      "lb_trace_inputs")
            else:
                a, b = payload
                buf = be.alloc(n, pooled=True)
                fn = lib.lb_trace_add if kind == "add" else lib.lb_trace_mul
                check(ctx, fn(ctx, node, a, b, C.c_void_p(values[a].ptr), C.c_void_p(values[b].ptr), n, mult, C.c_void_p(buf.ptr),
                              C.c_void_p(tables[kind].ptr), at[kind]), "lb_trace_" + kind)
            values.append(buf)
            at[kind] += n
        order = [k for k in ("add", "mul", "inputs") if rows_total[k]]
        device_tables = {k: (tables[k].ptr, rows_total[k], N_COLS[k]) for k in order}
        self.tables, self.values = tables, values  # the buffers must outlive the prove() call
        pie_meta = [(k, None) for k in order]  # table order for prove(); the rows live on the device (device_tables)
        return pie_meta, device_tables, values
