"""Device-side gen_trace (lb_trace_inputs / lb_trace_add / lb_trace_mul) against the host table builder, and proofs made
from device-generated tables against the oracle prover."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pie as piemod
from oracle import prover as oprover, verifier as overifier
from oracle.proof import from_bincode, to_bincode


@pytest.fixture(scope="module")
def be():
    from luminair_b200.backend import CudaBackend
    b = CudaBackend(0)
    yield b
    b.close()


def _both(be, n, seed):
    from luminair_b200.trace import DeviceGraphTrace
    rng = np.random.Generator(np.random.PCG64(seed))
    a = piemod.to_fixed(rng.uniform(-3.0, 3.0, n))
    b = piemod.to_fixed(rng.uniform(-3.0, 3.0, n))
    hg = piemod.GraphTrace()
    dg = DeviceGraphTrace(be)
    for g in (hg, dg):
        x, y = g.input(a), g.input(b)
        s = g.add(x, y)
        p = g.mul(s, x)      # negative products: floor division, remainder in [0, 4096)
        q = g.mul(p, p)
        g.add(q, y)          # final output: multiplicity 0
    return hg.finish()[0], dg.finish(), dg


@pytest.mark.parametrize("n", [1, 37, 4096, 100003])
def test_device_tables_equal_host_tables(be, n):
    host_pie, (meta, dev, values), dg = _both(be, n, seed=n)
    assert [k for k, _ in host_pie] == [k for k, _ in meta]
    for name, rows in host_pie:
        ptr, n_rows, n_cols = dev[name]
        assert (n_rows, n_cols) == rows.shape
        got = be.download(dg.tables[name]).reshape(n_rows, n_cols)
        assert np.array_equal(got, rows), name


def test_proof_from_device_generated_tables(be):
    from luminair_b200.prover import prove
    host_pie, (meta, dev, values), dg = _both(be, 300, seed=7)
    want = to_bincode(oprover.prove(host_pie))
    got = prove(meta, backend=be, device_tables=dev)
    assert got == want
    overifier.verify(from_bincode(got))


# ---- every operator (lb_trace_op): device tables vs the host builder, proofs from device-generated tables -----------------
def _compare(be, hg, dg, layouts=None):
    host_pie, host_pre = hg.finish(layouts=layouts)
    meta, dev, values = dg.finish(layouts if layouts is not None else hg.layouts)
    assert [k for k, _ in host_pie] == [k for k, _ in meta]
    for name, rows in host_pie:
        ptr, n_rows, n_cols = dev[name]
        assert (n_rows, n_cols) == rows.shape, name
        got = be.download(dg.tables[name]).reshape(n_rows, n_cols)
        assert np.array_equal(got, rows), name
    assert [c for c, _ in host_pre] == [c for c, _ in dg.preprocessed]
    for (_, a), (_, b) in zip(host_pre, dg.preprocessed):
        assert np.array_equal(a, b)
    for node, vals in enumerate(hg.values):  # the tensors themselves
        got = be.download(values[node], vals.size).view(np.int32)
        assert np.array_equal(got.astype(np.int64), vals), f"node {node}"
    return host_pie, host_pre, meta, dev


@pytest.mark.parametrize("n,seed", [(24, 3), (1000, 11), (40000, 5)])
def test_all_components_graph_tables(be, n, seed):
    from luminair_b200.trace import DeviceGraphTrace
    hg = piemod.build_all_components(piemod.GraphTrace(), n, seed)
    dg = piemod.build_all_components(DeviceGraphTrace(be), n, seed)
    host_pie, _, _, _ = _compare(be, hg, dg)
    assert len(host_pie) == 17


def test_gathers_broadcasts_and_signed_edge_cases(be):
    """Operands read through gather indices (repeats, permutations, broadcasts), negative dividends / divisors, equal
    operands of less_than, reductions over gathered groups, a node consumed several times."""
    from luminair_b200.trace import DeviceGraphTrace
    rng = np.random.Generator(np.random.PCG64(99))
    n = 777
    a = piemod.to_fixed(rng.uniform(-4.0, 4.0, n))
    b = piemod.to_fixed(rng.uniform(0.5, 4.0, n) * rng.choice([-1.0, 1.0], n))  # divisors: |b| >= 0.5 keeps 1/b in range
    a[a == 0] = -7
    b[:50] = a[:50]  # less_than with equal operands: diff = P
    perm = rng.permutation(n)
    rep = rng.integers(0, n, size=3 * n)
    perm3 = rng.permutation(3 * n)
    hg, dg = piemod.GraphTrace(), DeviceGraphTrace(be)
    for g in (hg, dg):
        x, y = g.input(a), g.input(b)
        r = g.rem((x, perm), y)
        q = g.recip((y, rep[:n]))
        lt = g.less_than(x, y)
        s = g.add((r, rep), (q, rep % n))            # 3n rows
        m = g.mul((lt, np.zeros(n, dtype=np.int64)), x)  # broadcast of element 0
        sr = g.sum_reduce((s, np.arange(3 * n)), 3)
        mr = g.max_reduce((s, perm3), 7)
        c = g.contiguous((mr, np.arange(3 * n // 7)[::-1]))
        g.add(sr, m)
        g.mul(c, c)
        g.max_reduce((s, perm3[:2330]), 2)
        g.sum_reduce((q, rep[:n]), 1)
    _compare(be, hg, dg)


def test_wide_graph_tables(be):
    from luminair_b200.trace import DeviceGraphTrace
    hg = piemod.build_wide(piemod.GraphTrace(), 12)
    dg = piemod.build_wide(DeviceGraphTrace(be), 12)
    _compare(be, hg, dg)


def test_lookup_input_outside_the_table_fails(be):
    from luminair_b200._lib import LuminairB200Error
    from luminair_b200.trace import DeviceGraphTrace
    dg = DeviceGraphTrace(be)
    dg.exp2(dg.input(piemod.to_fixed(np.linspace(-1.0, 1.0, 64))))
    with pytest.raises(LuminairB200Error, match="Value should fit in range"):
        dg.finish({"exp2": piemod.LookupLayout([(-4096, 100)])})


def test_lookup_layout_with_several_ranges(be):
    from luminair_b200.trace import DeviceGraphTrace
    v = np.concatenate([np.arange(-300, -200), np.arange(10, 90), np.arange(4000, 4100)]).astype(np.int64)
    layout = piemod.LookupLayout([(-320, -190), (0, 100), (3990, 4200)])
    dg = DeviceGraphTrace(be)
    out = dg.sin(dg.input(v))
    meta, dev, values = dg.finish({"sin": layout})
    want = piemod._apply_libm("sin", v)
    assert np.array_equal(be.download(values[out], v.size).view(np.int32).astype(np.int64), want)
    mult = be.download(dg.tables["sin_lookup"])
    want_mult = np.zeros(1 << layout.log_size, dtype=np.uint32)
    np.add.at(want_mult, layout.find_index(v), 1)
    assert np.array_equal(mult, want_mult)


def test_mlp_proof_from_device_generated_tables(be):
    """cfg-4 shape at reduced width: gen_trace on the device (Mul with broadcast operands, SumReduce, Add, Exp2 + LUT, Recip),
    proof bytes equal to the CPU restatement's proof of the host-built tables."""
    from luminair_b200.prover import prove
    from luminair_b200.trace import DeviceGraphTrace
    widths = (2, 8, 8, 1)
    hg = piemod.build_mlp(piemod.GraphTrace(), widths)
    dg = piemod.build_mlp(DeviceGraphTrace(be), widths)
    host_pie, host_pre, meta, dev = _compare(be, hg, dg)
    got = prove(meta, backend=be, device_tables=dev, preprocessed=dg.preprocessed)
    want = to_bincode(oprover.prove(host_pie, preprocessed=host_pre))
    assert got == want
    overifier.verify(from_bincode(got), preprocessed=[(cid, len(v).bit_length() - 1) for cid, v in overifier.sort_preprocessed(host_pre)])


def test_full_mlp_device_tables_prove_equals_host_tables_prove(be):
    from luminair_b200.prover import prove
    from luminair_b200.trace import DeviceGraphTrace
    hg = piemod.build_mlp(piemod.GraphTrace())
    dg = piemod.build_mlp(DeviceGraphTrace(be))
    host_pie, host_pre, meta, dev = _compare(be, hg, dg)
    assert prove(meta, backend=be, device_tables=dev, preprocessed=dg.preprocessed) == \
        prove(host_pie, backend=be, preprocessed=host_pre)


def test_replay_a_recorded_graph_with_new_inputs(be):
    """Record once, run per execution (the reference compiles the graph once and calls gen_trace per run): the second run
    uploads only the two input tensors and replays the emission; tables and lookup counters equal a fresh host build."""
    from luminair_b200.prover import prove
    from luminair_b200.trace import DeviceGraphTrace
    n = 600
    hg1 = piemod.build_all_components(piemod.GraphTrace(), n, 3)
    hg2 = piemod.build_all_components(piemod.GraphTrace(), n, 4)
    hg1.finish()
    hg2.finish()
    layouts = {}
    for name in hg1.layouts:  # settings that cover both executions
        (lo1, hi1), (lo2, hi2) = hg1.layouts[name].ranges[0], hg2.layouts[name].ranges[0]
        layouts[name] = piemod.LookupLayout([(min(lo1, lo2), max(hi1, hi2))])
    dg = piemod.build_all_components(DeviceGraphTrace(be), n, 3)
    hg1 = piemod.build_all_components(piemod.GraphTrace(), n, 3)
    _compare(be, hg1, dg, layouts)
    # second execution: new values for the two graph inputs (nodes 0 and 1)
    hg2 = piemod.build_all_components(piemod.GraphTrace(), n, 4)
    dg.set_input(0, hg2.values[0])
    dg.set_input(1, hg2.values[1])
    host_pie, host_pre, meta, dev = _compare(be, hg2, dg, layouts)
    assert prove(meta, backend=be, device_tables=dev, preprocessed=dg.preprocessed) == \
        prove(host_pie, backend=be, preprocessed=host_pre)
