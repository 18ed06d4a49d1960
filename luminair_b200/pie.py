"""Trace tables (the ``LuminairPie`` of crates/air/src/pie.rs:143-148) for element-wise graphs,
laid out the way ``LuminairGraph::gen_trace`` emits them (crates/graph/src/graph.rs:161-604):
one row per element with the columns of ``AddTraceTableRow`` (components/add/table.rs),
``MulTraceTableRow`` and ``InputsTraceTableRow``, multiplicities as the consumers dictate.

Host-side helper for benchmarks and tests: it only builds inputs for ``prove``.
"""
from __future__ import annotations

import numpy as np

P = (1 << 31) - 1
FP_SCALE = 1 << 12  # numerair Fixed<12>, crates/air/src/lib.rs:23


def to_fixed(x: np.ndarray) -> np.ndarray:
    """f32/f64 -> Fixed<12> raw value (round to nearest), as int64."""
    return np.round(np.asarray(x, dtype=np.float64) * FP_SCALE).astype(np.int64)


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
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << log_n
    a = to_fixed(rng.uniform(-0.5, 0.5, n))
    b = to_fixed(rng.uniform(-0.5, 0.5, n))
    return add_graph_pie(a, b)
