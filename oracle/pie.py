"""numpy restatement of ``LuminairGraph::gen_trace`` (oracle; TEST INFRASTRUCTURE ONLY): the trace tables (the
``LuminairPie`` of crates/air/src/pie.rs:143-148) of an operator graph over Fixed<12> tensors, laid out the way
/root/reference/crates/graph/src/graph.rs:161-604 emits them with the row emitters of crates/graph/src/op/prim.rs
(``process_trace`` of every operator).  It is the checker ``luminair_b200.trace.DeviceGraphTrace`` (gen_trace on the device) is
compared with, and the table builder of the oracle-side fixtures.  Self-contained: LUT columns are generated here with host
libm (``math``), independently of ``luminair_b200.lookups`` (tests/test_settings.py checks the two agree).
"""
from __future__ import annotations

import math

import numpy as np

# workload definitions (pure recorders against the GraphTrace interface; shared with the device trace)
from luminair_b200.workloads import build_add_graph, build_all_components, build_mlp, build_wide, synthetic_add_graph_inputs  # noqa: F401,E501

P = (1 << 31) - 1
FP_SCALE = 1 << 12  # numerair Fixed<12>, crates/air/src/lib.rs:23
RANGE_CHECK_BITS = 8


def _round_half_away(x) -> np.ndarray:
    """Rust f64::round (ties away from zero) - numerair Fixed::from_f64 restated as (v * 2^12).round(); parity unpinned."""
    x = np.asarray(x, dtype=np.float64)
    return (np.sign(x) * np.floor(np.abs(x) + 0.5)).astype(np.int64)


def to_fixed(x) -> np.ndarray:
    return _round_half_away(np.asarray(x, dtype=np.float64) * FP_SCALE)


def _apply_libm(name: str, raw) -> np.ndarray:
    """f(raw / 2^12) through the platform libm one element at a time (Rust's f64::sin/exp2/log2 are libm calls;
    preprocessed.rs:374,457,540), back to raw Fixed<12>; NaN / -inf -> 0."""
    fn = {"sin": math.sin, "exp2": lambda v: 2.0 ** v if not hasattr(math, "exp2") else math.exp2(v), "log2": math.log2}[name]
    vals = np.asarray(raw, dtype=np.int64).reshape(-1).tolist()
    out = np.empty(len(vals), dtype=np.float64)
    for i, v in enumerate(vals):
        try:
            out[i] = fn(v / FP_SCALE)
        except (ValueError, OverflowError):
            out[i] = 0.0
    out = np.nan_to_num(out, nan=0.0, neginf=0.0)
    return _round_half_away(out * FP_SCALE).reshape(np.shape(raw))


def _table(n, cols):
    out = np.empty((n, len(cols)), dtype=np.uint32)
    for j, c in enumerate(cols):
        out[:, j] = (np.broadcast_to(np.asarray(c, dtype=np.int64), (n,)) % P).astype(np.uint32)
    return out


def add_graph_pie(a_fixed: np.ndarray, b_fixed: np.ndarray):
    """c = a + b over n elements.  Node ids: a = 0, b = 1 (CopyToStwo inputs, op/prim.rs:72-84), add = 2
    (LuminairAdd::process_trace, op/prim.rs:967-1013).  -> [("add", [n,15]), ("inputs", [2n,7])]"""
    a = np.asarray(a_fixed, dtype=np.int64).reshape(-1)
    b = np.asarray(b_fixed, dtype=np.int64).reshape(-1)
    n = a.size
    idx = np.arange(n, dtype=np.int64)
    last = (idx == n - 1).astype(np.int64)
    add = _table(n, [2, 0, 1, idx, last, 2, 0, 1, idx + 1, a, b, a + b, -1, -1, 0])
    inp = np.concatenate([_table(n, [0, idx, last, 0, idx + 1, a, 1]), _table(n, [1, idx, last, 1, idx + 1, b, 1])])
    return [("add", add), ("inputs", inp)]


def synthetic_add_graph_pie(log_n: int, seed: int = 42):
    """BASELINE cfg 3: 2^log_n-element a + b with f32 uniform(-0.5, 0.5) inputs, PCG64(seed)."""
    return add_graph_pie(*synthetic_add_graph_inputs(log_n, seed))


class LookupLayout:
    """crates/air/src/preprocessed.rs:41-116: sorted, disjoint value ranges [(lo, hi)] (inclusive raw Fixed<12>
    values); entry i of the table is the i-th value of the concatenated ranges."""

    def __init__(self, ranges):
        self.ranges = sorted((int(lo), int(hi)) for lo, hi in ranges)
        count = sum(hi - lo + 1 for lo, hi in self.ranges)
        # calculate_log_size, crates/air/src/utils.rs:22-27
        self.log_size = max(((count + 15) >> 4) - 1, 0).bit_length() + 4
        self.values = np.concatenate([np.arange(lo, hi + 1, dtype=np.int64) for lo, hi in self.ranges])

    @staticmethod
    def covering(values, pad: int = 0):
        """One range covering the given raw values (the reference derives ranges from a calibration run,
        crates/graph/src/graph.rs gen_circuit_settings)."""
        v = np.asarray(values, dtype=np.int64)
        return LookupLayout([(int(v.min()) - pad, int(v.max()) + pad)])

    def find_index(self, target: np.ndarray) -> np.ndarray:
        idx = np.searchsorted(self.values, target)
        if np.any(idx >= self.values.size) or np.any(self.values[np.minimum(idx, self.values.size - 1)] != target):
            raise ValueError("Value should fit in range.")
        return idx


def lut_columns(name: str, layout: LookupLayout):
    """Preprocessed LUT columns as the reference generates them on the host (preprocessed.rs:351-383 sin,
    :436-462 exp2, log2 likewise): column 0 = the input values, column 1 = f(value) through f64 and
    Fixed::from_f64, zero beyond the enumerated values.  -> [(id, values[2^log_size])]"""
    n = 1 << layout.log_size
    c0 = np.zeros(n, dtype=np.int64)
    c1 = np.zeros(n, dtype=np.int64)
    c0[: layout.values.size] = layout.values
    c1[: layout.values.size] = _apply_libm(name, layout.values)
    return [(f"{name}_lut_0", (c0 % P).astype(np.uint32)), (f"{name}_lut_1", (c1 % P).astype(np.uint32))]


def range_check_column(n_bits: int = RANGE_CHECK_BITS):
    """preprocessed.rs:233-248,289-307: the enumeration 0 .. 2^n_bits - 1 (one segment)."""
    log_size = max(n_bits, 4)
    col = np.zeros(1 << log_size, dtype=np.uint32)
    col[: 1 << n_bits] = np.arange(1 << n_bits, dtype=np.uint32)
    return [(f"range_check_{n_bits}_column_0", col)]


class GraphTrace:
    """Builds the ``LuminairPie`` trace tables of an operator graph over Fixed<12> tensors the way
    ``LuminairGraph::gen_trace`` does (crates/graph/src/graph.rs:161-604 with the row emitters of
    crates/graph/src/op/prim.rs): every operator node appends one row per output element (per step for the
    reductions) to its component's table; producers yield each element with multiplicity = number of rows that
    read it, consumers take it with -1; LUT operators count their lookups in the table's multiplicity column.

    Operands are (node, gather index) pairs so broadcasts / expands are explicit: ``idx[i]`` is the element of
    the input tensor that output row ``i`` reads (default: the identity).
    """

    def __init__(self):
        self.values = []     # node -> raw int64 tensor (flattened)
        self.uses = []       # node -> per-element consumer count
        self.ops = []        # (kind, node, dict)
        self.lut_inputs = {"sin": [], "exp2": [], "log2": []}
        self.limbs = []

    # -- nodes ------------------------------------------------------------------------------
    def _node(self, vals):
        self.values.append(np.asarray(vals, dtype=np.int64).reshape(-1))
        self.uses.append(np.zeros(self.values[-1].size, dtype=np.int64))
        return len(self.values) - 1

    def _read(self, operand, n=None):
        node, idx = operand if isinstance(operand, tuple) else (operand, None)
        if idx is None:
            idx = np.arange(self.values[node].size if n is None else n, dtype=np.int64)
        idx = np.asarray(idx, dtype=np.int64)
        np.add.at(self.uses[node], idx, 1)
        return node, self.values[node][idx]

    def input(self, raw):
        """CopyToStwo of a graph input / initializer (op/prim.rs:72-84) -> Inputs table."""
        node = self._node(raw)
        self.ops.append(("inputs", node, {}))
        return node

    def _binary(self, kind, a, b, fn):
        ia, va = self._read(a)
        ib, vb = self._read(b)
        out, extra = fn(va, vb)
        node = self._node(out)
        self.ops.append((kind, node, dict(lhs_id=ia, rhs_id=ib, lhs=va, rhs=vb, **extra)))
        return node

    def add(self, a, b):
        return self._binary("add", a, b, lambda x, y: (x + y, {}))

    def mul(self, a, b):
        def f(x, y):
            p = x * y
            q = p // FP_SCALE
            return q, {"rem": p - q * FP_SCALE}
        return self._binary("mul", a, b, f)

    def rem(self, a, b):
        def f(x, y):
            q = x // y
            r = x - q * y
            return r, {"quotient": q}
        return self._binary("rem", a, b, f)

    def less_than(self, a, b):
        def f(x, y):
            lt = x < y
            diff = np.where(lt, y - x, y - x + P)  # op/prim.rs:1209-1213 (two_pow_k = 2^31 - 1)
            d32 = diff.astype(np.uint32)
            limbs = [((d32 >> (8 * k)) & 0xFF).astype(np.int64) for k in range(4)]
            self.limbs.extend(limbs)
            return np.where(lt, FP_SCALE, 0).astype(np.int64), {"borrow": (~lt).astype(np.int64), "diff": diff % P, "limbs": limbs}
        return self._binary("less_than", a, b, f)

    def _unary(self, kind, a, fn):
        ia, va = self._read(a)
        out, extra = fn(va)
        node = self._node(out)
        self.ops.append((kind, node, dict(input_id=ia, input=va, **extra)))
        return node

    def recip(self, a):
        def f(x):
            q = (FP_SCALE * FP_SCALE) // x
            return q, {"rem": FP_SCALE * FP_SCALE - q * x}
        return self._unary("recip", a, f)

    def sqrt(self, a):
        def f(x):
            t = x * FP_SCALE
            s = np.floor(np.sqrt(t.astype(np.float64))).astype(np.int64)
            s = np.where(s * s > t, s - 1, s)
            s = np.where((s + 1) * (s + 1) <= t, s + 1, s)
            return s, {"rem": t - s * s}
        return self._unary("sqrt", a, f)

    def _lut(self, name, a):
        def f(x):
            self.lut_inputs[name].append(x)
            return _apply_libm(name, x), {}
        return self._unary(name, a, f)

    def sin(self, a):
        return self._lut("sin", a)

    def exp2(self, a):
        return self._lut("exp2", a)

    def log2(self, a):
        return self._lut("log2", a)

    def contiguous(self, a):
        return self._unary("contiguous", a, lambda x: (x.copy(), {}))

    def _reduce(self, kind, a, group):
        """a: operand whose gathered tensor is [n_out, group] row-major (op/prim.rs:1517-1565)."""
        ia, va = self._read(a)
        xs = va.reshape(-1, group)
        if kind == "sum_reduce":
            nxt = np.cumsum(xs, axis=1)
            extra = {"acc": nxt - xs, "next": nxt}
        else:
            run = np.maximum.accumulate(xs, axis=1)
            prev = np.concatenate([np.zeros((xs.shape[0], 1), dtype=np.int64), run[:, :-1]], axis=1)
            is_max = np.concatenate([np.ones((xs.shape[0], 1), dtype=np.int64), (xs[:, 1:] > run[:, :-1]).astype(np.int64)], axis=1)
            extra = {"acc": prev, "next": run, "is_max": is_max}
        node = self._node(extra["next"][:, -1])
        self.ops.append((kind, node, dict(input_id=ia, input=xs, **extra)))
        return node

    def sum_reduce(self, a, group):
        return self._reduce("sum_reduce", a, group)

    def max_reduce(self, a, group):
        return self._reduce("max_reduce", a, group)

    # -- tables ------------------------------------------------------------------------------
    def finish(self, lut_pad: int = 0, layouts=None):
        """-> (pie, preprocessed): trace tables in claim-slot order (graph.rs:502-593) and the LUT columns in
        ``lookups_to_preprocessed_column`` order (preprocessed.rs:181-206).  Nodes nobody reads are the graph's
        final outputs (multiplicity 0).  ``layouts``: {lut name: LookupLayout} of given circuit settings (default: one range
        covering the values seen, as a calibration run would produce)."""
        rows = {}

        def emit(kind, cols, n):
            rows.setdefault(kind, []).append(_table(n, cols))

        for kind, node, d in self.ops:
            out = self.values[node]
            mult = self.uses[node]
            n = out.size
            idx = np.arange(n, dtype=np.int64)
            last = (idx == n - 1).astype(np.int64)
            if kind == "inputs":
                emit(kind, [node, idx, last, node, idx + 1, out, mult], n)
            elif kind in ("add", "mul", "rem", "less_than"):
                head = [node, d["lhs_id"], d["rhs_id"], idx, last, node, d["lhs_id"], d["rhs_id"], idx + 1]
                if kind == "add":
                    emit(kind, head + [d["lhs"], d["rhs"], out, -1, -1, mult], n)
                elif kind == "mul":
                    emit(kind, head + [d["lhs"], d["rhs"], out, d["rem"], -1, -1, mult], n)
                elif kind == "rem":
                    emit(kind, head + [d["lhs"], d["rhs"], out, d["quotient"], -1, -1, mult], n)
                else:
                    emit(kind, head + [d["lhs"], d["rhs"], out, d["diff"], d["borrow"]] + d["limbs"] + [-1, -1, mult, 1], n)
            elif kind in ("recip", "sqrt"):
                emit(kind, [node, d["input_id"], idx, last, node, d["input_id"], idx + 1, d["input"], out, d["rem"], FP_SCALE, -1, mult], n)
            elif kind in ("sin", "exp2", "log2"):
                emit(kind, [node, d["input_id"], idx, last, node, d["input_id"], idx + 1, d["input"], out, -1, mult, 1], n)
            elif kind == "contiguous":
                emit(kind, [node, d["input_id"], idx, last, node, d["input_id"], idx + 1, d["input"], out, -1, mult], n)
            else:  # reductions: one row per (output element, step)
                m, k = d["input"].shape
                out_idx = np.repeat(np.arange(m, dtype=np.int64), k)
                step = np.tile(np.arange(k, dtype=np.int64), m)
                last_idx = (out_idx == m - 1).astype(np.int64)
                last_step = (step == k - 1).astype(np.int64)
                # the finished value is yielded once per output element, on its last step
                cols = [node, d["input_id"], out_idx, last_idx, node, d["input_id"], out_idx + 1, d["input"].reshape(-1),
                        np.repeat(out, k), d["acc"].reshape(-1), d["next"].reshape(-1), last_step]
                if kind == "max_reduce":
                    cols.append(d["is_max"].reshape(-1))
                cols += [-1, np.repeat(mult, k) * last_step]
                emit(kind, cols, m * k)

        preprocessed = []
        lut_tables = {}
        self.layouts = {}  # what the reference's calibration pass (gen_circuit_settings) would put in the settings
        for name in ("sin", "exp2", "log2"):
            if not self.lut_inputs[name]:
                continue
            x = np.concatenate(self.lut_inputs[name])
            layout = self.layouts[name] = (layouts or {}).get(name) or LookupLayout.covering(x, lut_pad)
            preprocessed += lut_columns(name, layout)
            mult = np.zeros(1 << layout.log_size, dtype=np.int64)
            np.add.at(mult, layout.find_index(x), 1)
            lut_tables[name + "_lookup"] = mult.reshape(-1, 1)
        if self.limbs:
            preprocessed += range_check_column()
            mult = np.zeros(1 << max(RANGE_CHECK_BITS, 4), dtype=np.int64)
            np.add.at(mult, np.concatenate(self.limbs), 1)
            lut_tables["range_check_lookup"] = mult.reshape(-1, 1)

        order = ["add", "mul", "recip", "sin", "sin_lookup", "sum_reduce", "max_reduce", "sqrt", "rem", "exp2", "exp2_lookup",
                 "log2", "log2_lookup", "less_than", "range_check_lookup", "inputs", "contiguous"]
        pie = []
        for kind in order:
            if kind in rows:
                pie.append((kind, np.concatenate(rows[kind])))
            elif kind in lut_tables:
                pie.append((kind, (lut_tables[kind] % P).astype(np.uint32)))
        return pie, preprocessed


def all_components_graph(n: int = 24, seed: int = 3):
    """A graph that touches all 17 components (tests)."""
    return build_all_components(GraphTrace(), n, seed).finish()


def mlp_graph(widths=(2, 64, 64, 1), x=(15.0, 0.5), seed: int = 7, scale: float = 0.3):
    """BASELINE cfg 4 shape: see luminair_b200.workloads.build_mlp."""
    return build_mlp(GraphTrace(), widths, x, seed, scale).finish()


def wide_graph(log_n: int, seed: int = 64):
    """The headline trace shape ("2^20 x 64"): see luminair_b200.workloads.build_wide."""
    pie, pre = build_wide(GraphTrace(), log_n, seed).finish()
    assert not pre
    return pie
