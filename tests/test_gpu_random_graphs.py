"""Randomised operator graphs: for every seed a different mix of the 17 components, node counts, tensor sizes
(ragged: not powers of two), broadcast reads and LUT ranges.  The CUDA prover's bytes must equal the oracle prover's
and pass the oracle verifier.  Sizes stay small so the numpy oracle finishes in seconds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pie as piemod
from oracle import prover as oprover, verifier as overifier
from oracle.proof import from_bincode, to_bincode
from test_gpu_prover import _assert_same_proof, _oracle_transcript


@pytest.fixture(scope="module")
def be():
    from luminair_b200.backend import CudaBackend
    b = CudaBackend(0)
    yield b
    b.close()


class _Both:
    """Records every node on the host builder and on a second builder with the same interface (same node ids)."""

    def __init__(self, host, other):
        self.host, self.other, self.values = host, other, host.values

    def __getattr__(self, name):
        def call(*args):
            node = getattr(self.host, name)(*args)
            assert getattr(self.other, name)(*args) == node
            return node
        return call


def random_graph(seed: int, builder=None, n_range=(5, 90)):
    rng = np.random.Generator(np.random.PCG64(seed))
    g = builder or piemod.GraphTrace()
    n = int(rng.integers(*n_range))
    pos = [g.input(piemod.to_fixed(rng.uniform(0.3, 3.0, n)))]   # strictly positive tensors (recip, sqrt, log2, rem divisor)
    any_ = [g.input(piemod.to_fixed(rng.uniform(-2.0, 2.0, n)))]  # any sign
    for _ in range(int(rng.integers(3, 10))):
        op = rng.choice(["add", "mul", "recip", "sqrt", "rem", "exp2", "log2", "sin", "less_than", "sum_reduce", "max_reduce",
                         "contiguous", "broadcast_mul"])
        a = any_[int(rng.integers(len(any_)))]
        b = any_[int(rng.integers(len(any_)))]
        p = pos[int(rng.integers(len(pos)))]
        if g.values[a].size != g.values[b].size:
            b = a
        if op == "add":
            any_.append(g.add(a, b))
        elif op == "mul":
            any_.append(g.mul(a, b))
        elif op == "recip":
            pos.append(g.recip(p))
        elif op == "sqrt":
            pos.append(g.sqrt(p))
        elif op == "rem":
            q = pos[int(rng.integers(len(pos)))]
            if g.values[q].size == g.values[p].size:
                any_.append(g.rem(p, q))
        elif op == "exp2":
            pos.append(g.exp2(a))
        elif op == "log2":
            any_.append(g.log2(p))
        elif op == "sin":
            any_.append(g.sin(a))
        elif op == "less_than":
            any_.append(g.less_than(a, b))
        elif op in ("sum_reduce", "max_reduce"):
            size = g.values[a].size
            k = int(rng.choice([d for d in (1, 2, 3, 4, 5, 6) if size % d == 0]))
            any_.append(getattr(g, op)(a, k))
        elif op == "contiguous":
            any_.append(g.contiguous(a))
        else:  # a broadcast read: every element of `a` against a cyclic view of `p`
            size = g.values[a].size
            idx = np.arange(size, dtype=np.int64) % g.values[p].size
            any_.append(g.mul(a, (p, idx)))
    return g if builder is not None else g.finish()


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106, 107, 108])
def test_random_graph_proof_bytes(be, seed):
    from luminair_b200.prover import prove
    pie, pre = random_graph(seed)
    lp, digests = _oracle_transcript(lambda: oprover.prove(pie, preprocessed=pre))
    got = prove(pie, backend=be, preprocessed=pre)
    _assert_same_proof(be, got, to_bincode(lp), digests)
    overifier.verify(from_bincode(got), preprocessed=[(cid, len(v).bit_length() - 1) for cid, v in pre])


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106, 107, 108, 201, 202, 203, 204])
def test_random_graph_device_gen_trace(be, seed):
    """The same random graphs with gen_trace on the device (lb_trace_op): every table, lookup multiplicity table and node tensor
    equals the host builder's, and the proof from the device-generated tables has the same bytes."""
    from luminair_b200.prover import prove
    from luminair_b200.trace import DeviceGraphTrace
    from test_gpu_trace import _compare
    hg, dg = piemod.GraphTrace(), DeviceGraphTrace(be)
    random_graph(seed, _Both(hg, dg))
    host_pie, host_pre, meta, dev = _compare(be, hg, dg)
    assert prove(meta, backend=be, device_tables=dev, preprocessed=dg.preprocessed) == \
        prove(host_pie, backend=be, preprocessed=host_pre)


@pytest.mark.parametrize("seed,n_range", [(301, (1 << 10, 1 << 12)), (302, (1 << 12, 1 << 14)), (303, (1 << 14, 1 << 16)),
                                          (304, (1 << 15, 1 << 17)), (305, (1 << 16, 1 << 18)), (306, (1 << 13, 1 << 17))])
def test_random_graph_at_size_equals_cpu_prover(be, seed, n_range):
    """Mid-size random graphs (ragged tensor lengths up to 2^18, tables up to 2^20 rows): too large for the numpy oracle inside a
    test, so the checker is the compiled CPU prover (oracle/c/cpu_prover, pinned to the numpy oracle and the reference's committed
    proof by tests/test_cpu_prover.py)."""
    from luminair_b200.prover import prove
    from oracle import cpu_prover as cp
    pie, pre = random_graph(seed, n_range=n_range)
    want = cp.prove(pie, preprocessed=pre)
    got = prove(pie, backend=be, preprocessed=pre)
    assert got == want
