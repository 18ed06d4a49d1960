"""Device-side gen_trace (lb_trace_inputs / lb_trace_add / lb_trace_mul) against the host table builder, and proofs made
from device-generated tables against the oracle prover."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from luminair_b200 import pie as piemod
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
