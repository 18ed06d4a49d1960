"""Workload graphs of the BASELINE configs, written against the graph-recording interface (``input / add / mul / rem /
less_than / recip / sqrt / sin / exp2 / log2 / contiguous / sum_reduce / max_reduce``) that both ``trace.DeviceGraphTrace``
(gen_trace on the device) and the numpy checker of the tests implement.  Only definitions of *what* is proved:
no trace generation happens here."""
from __future__ import annotations

import numpy as np

from .lookups import FP_SCALE, to_fixed


def synthetic_add_graph_inputs(log_n: int, seed: int = 42):
    """The two input tensors of BASELINE cfg 3 as raw Fixed<12> values: f32 uniform(-0.5, 0.5), PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << log_n
    a = to_fixed(rng.uniform(-0.5, 0.5, n))
    b = to_fixed(rng.uniform(-0.5, 0.5, n))
    return a, b


# The Exp2 lookup range a calibration run (the reference's gen_circuit_settings) of the default cfg-4 MLP below yields: the raw
# Fixed<12> inputs its Exp2 nodes see (tests/test_settings.py checks it against the values the numpy checker records)
MLP_EXP2_RANGE = (-57162, 54948)


def build_add_graph(g, a_fixed, b_fixed):
    """BASELINE cfg 3: c = a + b over n elements (nodes: a = 0, b = 1, add = 2) -> Add table n rows, Inputs table 2n rows."""
    a = g.input(a_fixed)
    b = g.input(b_fixed)
    g.add(a, b)
    return g


def build_all_components(g, n: int = 24, seed: int = 3):
    """Record the all-components graph on `g` (a GraphTrace or a trace.DeviceGraphTrace: same interface)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = g.input(to_fixed(rng.uniform(0.25, 2.0, n)))
    y = g.input(to_fixed(rng.uniform(-1.0, 1.0, n)))
    s = g.add(x, y)
    p = g.mul(s, x)
    r = g.recip(x)
    q = g.sqrt(x)
    e = g.exp2(y)
    l = g.log2(x)
    t = g.sin(p)
    m = g.rem(x, q)
    c = g.less_than(y, r)
    k = 4
    sr = g.sum_reduce(e, k)
    mr = g.max_reduce(l, k)
    g.contiguous(t)
    g.add(m, c)
    g.mul(sr, mr)
    return g


def build_mlp(g, widths=(2, 64, 64, 1), x=(15.0, 0.5), seed: int = 7, scale: float = 0.3):
    """BASELINE cfg 4 shape (examples/black-schole-nn/src/main.rs: Linear 2-64-64-1 with tanh between, input
    [15.0, 0.5]); synthetic weights uniform(-scale, scale), PCG64(seed) (the reference's weights are git-ignored).
    Each Linear is Mul over the expanded [out, in] operands + SumReduce + bias Add; tanh(z) lowers the way luminal
    does it: 2 * sigmoid(2z) - 1 with sigmoid(v) = 1 / (1 + exp2(-v * log2 e))  ->  Mul, Exp2, Add, Recip."""
    rng = np.random.Generator(np.random.PCG64(seed))
    act = g.input(to_fixed(np.asarray(x, dtype=np.float64)))
    n_layers = len(widths) - 1
    for li in range(n_layers):
        d_in, d_out = widths[li], widths[li + 1]
        w = g.input(to_fixed(rng.uniform(-scale, scale, d_out * d_in)))
        b = g.input(to_fixed(rng.uniform(-scale, scale, d_out)))
        prod = g.mul(w, (act, np.tile(np.arange(d_in, dtype=np.int64), d_out)))
        z = g.add(g.sum_reduce(prod, d_in), b) if d_in > 1 else g.add(prod, b)
        if li == n_layers - 1:
            act = z
            break
        c_m2log2e = g.input(np.full(d_out, int(round(-2.0 * np.log2(np.e) * FP_SCALE))))
        one = g.input(np.full(d_out, FP_SCALE))
        two = g.input(np.full(d_out, 2 * FP_SCALE))
        neg_one = g.input(np.full(d_out, -FP_SCALE))
        e = g.exp2(g.mul(z, c_m2log2e))
        sig = g.recip(g.add(e, one))
        act = g.add(g.mul(sig, two), neg_one)
    return g


def build_wide(g, log_n: int, seed: int = 64):
    """The headline trace shape of BASELINE.json ("2^20 x 64"): four 2^log_n-row operator tables over the same two input
    tensors - Add (15 columns), Mul (16), Rem (16), SumReduce with groups of one (14) = 61 main-trace columns of 2^log_n rows,
    beside the Inputs table (7 columns, 2^(log_n+1) rows).  Inputs are positive Fixed<12> values (Rem needs a non-zero
    divisor), uniform(0.25, 2), PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << log_n
    a = g.input(to_fixed(rng.uniform(0.25, 2.0, n)))
    b = g.input(to_fixed(rng.uniform(0.25, 2.0, n)))
    g.add(a, b)
    g.mul(a, b)
    g.rem(a, b)
    g.sum_reduce(a, 1)
    return g
