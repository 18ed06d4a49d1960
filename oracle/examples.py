"""Trace tables for the BASELINE configs (oracle; test infrastructure only).

``simple_pie`` follows examples/simple/src/main.rs:15-22 + the row emitters in
crates/graph/src/op/prim.rs (LuminairAdd::process_trace :967-1013, LuminairMul,
CopyToStwo :72-84) with the node ids of ui/demo/public/graph.dot.
``synthetic_pie`` builds satisfiable Mul-shaped tables of any size (BASELINE cfg 3).
"""
from __future__ import annotations

import numpy as np

from .fields import P, U64

SCALE = 4096  # Fixed<12>


def _m(v):
    return v % P


def add_rows(node, lid, rid, l, r, o, lm, rm, om):
    n = len(l)
    return [[node, lid, rid, i, 1 if i == n - 1 else 0, node, lid, rid, i + 1,
             _m(l[i]), _m(r[i]), _m(o[i]), _m(lm), _m(rm), _m(om)] for i in range(n)]


def mul_rows(node, lid, rid, l, r, o, rem, lm, rm, om):
    n = len(l)
    return [[node, lid, rid, i, 1 if i == n - 1 else 0, node, lid, rid, i + 1,
             _m(l[i]), _m(r[i]), _m(o[i]), _m(rem[i]), _m(lm), _m(rm), _m(om)] for i in range(n)]


def inputs_rows(node, vals, mult):
    n = len(vals)
    return [[node, i, 1 if i == n - 1 else 0, node, i + 1, _m(vals[i]), _m(mult)] for i in range(n)]


def fixed_mul(x, y):
    """numerair Fixed<12> mul -> (q, r) with x*y = q*4096 + r, floor division (r >= 0)."""
    p = x * y
    return p // SCALE, p % SCALE


def simple_values():
    a = [1, 2, 3, 4]
    b = [10, 20, 30, 40]
    w = [-1] * 4
    fa, fb, fw = ([x * SCALE for x in v] for v in (a, b, w))
    c, crem = zip(*[fixed_mul(x, y) for x, y in zip(fa, fb)])
    d = [x + y for x, y in zip(c, fw)]
    e, erem = zip(*[fixed_mul(x, y) for x, y in zip(c, d)])
    return fa, fb, fw, list(c), list(crem), d, list(e), list(erem)


def simple_pie(era: str = "current"):
    """era="artifact": schema of the committed UI proof (no Inputs component; inputs that
    come from initializers carry multiplicity 0).  era="current": schema of the tree at
    /root/reference (Inputs component yields every CopyToStwo value)."""
    fa, fb, fw, c, crem, d, e, erem = simple_values()
    if era == "artifact":
        add = add_rows(4, 3, 8, c, fw, d, -1, 0, 1)
        mul = mul_rows(3, 6, 7, fa, fb, c, crem, 0, 0, 2) + mul_rows(5, 3, 4, c, d, e, erem, -1, -1, 0)
        return [("add", np.array(add, dtype=U64)), ("mul", np.array(mul, dtype=U64))]
    add = add_rows(4, 3, 8, c, fw, d, -1, -1, 1)
    mul = mul_rows(3, 6, 7, fa, fb, c, crem, -1, -1, 2) + mul_rows(5, 3, 4, c, d, e, erem, -1, -1, 0)
    inp = inputs_rows(6, fa, 1) + inputs_rows(7, fb, 1) + inputs_rows(8, fw, 1)
    return [("add", np.array(add, dtype=U64)), ("mul", np.array(mul, dtype=U64)), ("inputs", np.array(inp, dtype=U64))]


def synthetic_mul_table(log_rows: int, seed: int, node_id: int = 3) -> np.ndarray:
    """One Mul node over 2^log_rows elements with random Fixed<12> operands in (-0.5, 0.5);
    every multiplicity 0 except a self-cancelling pair so the LogUp sum is zero:
    lhs consumed once (-1) and yielded once by `out` of the same value?  No - to keep the
    table self-contained we set lhs == out-of-previous semantics aside and simply use
    lhs_mult = rhs_mult = out_mult = 0 (constraints hold; claimed sum 0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << log_rows
    x = np.round(rng.uniform(-0.5, 0.5, n) * SCALE).astype(np.int64)
    y = np.round(rng.uniform(-0.5, 0.5, n) * SCALE).astype(np.int64)
    p = x * y
    q = p // SCALE
    r = p - q * SCALE
    rows = np.zeros((n, 16), dtype=np.int64)
    idx = np.arange(n)
    rows[:, 0] = node_id
    rows[:, 1] = node_id + 3
    rows[:, 2] = node_id + 4
    rows[:, 3] = idx
    rows[:, 4] = (idx == n - 1)
    rows[:, 5] = node_id
    rows[:, 6] = node_id + 3
    rows[:, 7] = node_id + 4
    rows[:, 8] = idx + 1
    rows[:, 9] = x
    rows[:, 10] = y
    rows[:, 11] = q
    rows[:, 12] = r
    return (rows % P).astype(U64)


def graph_pie(log_n: int, seed: int = 42, with_mul: bool = True):
    """Trace tables of the element-wise graph  c = a + b ; e = c * a  over n = 2^log_n elements
    (BASELINE cfg 3 shape when with_mul=False: Add 2^log_n rows, Inputs 2^(log_n+1) rows), built the
    way the operators emit rows (crates/graph/src/op/prim.rs:72-84, 967-1013) with consistent LogUp
    multiplicities, so the proof verifies (claimed sums cancel).  Node ids: a=0, b=1, add=2, mul=3.
    Inputs: Fixed<12> of uniform(-0.5, 0.5), PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << log_n
    a = np.round(rng.uniform(-0.5, 0.5, n) * SCALE).astype(np.int64)
    b = np.round(rng.uniform(-0.5, 0.5, n) * SCALE).astype(np.int64)
    c = a + b
    idx = np.arange(n, dtype=np.int64)
    last = (idx == n - 1).astype(np.int64)

    def table(cols):
        return (np.stack([np.broadcast_to(np.asarray(x, dtype=np.int64), (n,)) for x in cols], axis=1) % P).astype(U64)

    add = table([2, 0, 1, idx, last, 2, 0, 1, idx + 1, a, b, c, -1, -1, 1 if with_mul else 0])
    a_mult = 2 if with_mul else 1
    inp = np.concatenate([table([0, idx, last, 0, idx + 1, a, a_mult]), table([1, idx, last, 1, idx + 1, b, 1])])
    pie = [("add", add)]
    if with_mul:
        p = c * a
        q = p // SCALE
        r = p - q * SCALE
        mul = table([3, 2, 0, idx, last, 3, 2, 0, idx + 1, c, a, q, r, -1, -1, 0])
        pie.append(("mul", mul))
    pie.append(("inputs", inp))
    return pie


def reduce_pie(log_n: int, log_group: int = 2, seed: int = 1):
    """Graph with the reduction / copy operators:  x (input, node 0, n = 2^log_n elements) feeds
    SumReduce (node 1), MaxReduce (node 2) over groups of 2^log_group elements, and Contiguous (node 3).
    Rows as emitted by LuminairSumReduce / LuminairMaxReduce / LuminairContiguous::process_trace
    (crates/graph/src/op/prim.rs:1517-1565 and siblings): one row per (output element, step); all three
    outputs are final (multiplicity 0), x is consumed three times."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << log_n
    k = 1 << log_group
    m = n // k
    x = np.round(rng.uniform(-0.5, 0.5, n) * SCALE).astype(np.int64)
    xs = x.reshape(m, k)
    out_idx = np.repeat(np.arange(m, dtype=np.int64), k)
    step = np.tile(np.arange(k, dtype=np.int64), m)
    last_idx = (out_idx == m - 1).astype(np.int64)
    last_step = (step == k - 1).astype(np.int64)

    def table(cols):
        return (np.stack([np.broadcast_to(np.asarray(c, dtype=np.int64), (n,)) for c in cols], axis=1) % P).astype(U64)

    # sum reduce
    nxt = np.cumsum(xs, axis=1)
    acc = nxt - xs
    out = np.repeat(nxt[:, -1], k)
    sum_t = table([1, 0, out_idx, last_idx, 1, 0, out_idx + 1, x, out, acc.reshape(-1), nxt.reshape(-1), last_step, -1, 0])
    # max reduce: running max starts at the first element (is_max = 1 on step 0)
    run = np.maximum.accumulate(xs, axis=1)
    prev = np.concatenate([np.zeros((m, 1), dtype=np.int64), run[:, :-1]], axis=1)
    is_max = np.concatenate([np.ones((m, 1), dtype=np.int64), (xs[:, 1:] > run[:, :-1]).astype(np.int64)], axis=1)
    mout = np.repeat(run[:, -1], k)
    max_t = table([2, 0, out_idx, last_idx, 2, 0, out_idx + 1, x, mout, prev.reshape(-1), run.reshape(-1), last_step,
                   is_max.reshape(-1), -1, 0])
    idx = np.arange(n, dtype=np.int64)
    last = (idx == n - 1).astype(np.int64)
    inp = table([0, idx, last, 0, idx + 1, x, 3])
    cont = table([3, 0, idx, last, 3, 0, idx + 1, x, x, -1, 0])
    return [("sum_reduce", sum_t), ("max_reduce", max_t), ("inputs", inp), ("contiguous", cont)]
