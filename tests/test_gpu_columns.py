"""The small backend-trait methods (ColumnOps::bit_reverse_column, PolyOps::new_canonical_ordered, FieldOps::batch_inverse,
AccumulationOps::accumulate) through the C ABI against the CPU restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import air as oair
from oracle.circle import bit_reverse_indices
from oracle.fields import P, QM31


@pytest.fixture(scope="module")
def be():
    from luminair_b200.backend import CudaBackend
    b = CudaBackend(0)
    yield b
    b.close()


def _rand(rng, n):
    return rng.integers(0, P, size=n, dtype=np.uint64).astype(np.uint32)


@pytest.mark.parametrize("log", [0, 1, 5, 12, 20])
def test_bit_reverse_column(be, log):
    rng = np.random.Generator(np.random.PCG64(log))
    v = _rand(rng, 1 << log)
    buf = be.upload(v)
    be.bit_reverse_column(buf.ptr, log)
    assert np.array_equal(be.download(buf), v[bit_reverse_indices(log)])


@pytest.mark.parametrize("log", [1, 4, 11, 18])
def test_new_canonical_ordered(be, log):
    rng = np.random.Generator(np.random.PCG64(100 + log))
    v = _rand(rng, 1 << log)
    src, dst = be.upload(v), be.alloc(1 << log)
    be.new_canonical_ordered(src.ptr, dst.ptr, log)
    want = np.empty_like(v)
    want[oair.coset_order_storage_perm(log)] = v
    assert np.array_equal(be.download(dst), want)


@pytest.mark.parametrize("n", [1, 3, 4, 1000, (1 << 18) + 5])
def test_batch_inverse_base_field(be, n):
    rng = np.random.Generator(np.random.PCG64(n))
    v = rng.integers(1, P, size=n, dtype=np.uint64)
    buf = be.upload(v.astype(np.uint32))
    be.batch_inverse(buf.ptr, buf.ptr, n)  # in place
    got = be.download(buf).astype(np.uint64)
    assert np.all((got * v) % P == 1)
    assert np.all(got < P)


@pytest.mark.parametrize("n", [1, 6, 4097])
def test_batch_inverse_secure_field(be, n):
    rng = np.random.Generator(np.random.PCG64(7 * n))
    v = rng.integers(0, P, size=(4, n), dtype=np.uint64)
    src = be.upload(v.astype(np.uint32).reshape(-1))
    dst = be.alloc(4 * n)
    be.batch_inverse_secure([src.at(c * n) for c in range(4)], [dst.at(c * n) for c in range(4)], n)
    got = be.download(dst).reshape(4, n)
    want = QM31.from_array(v.T).inv().to_array()  # [n, 4]
    assert np.array_equal(got.T, want)


def test_batch_inverse_of_zero_fails(be):
    from luminair_b200._lib import LuminairB200Error
    buf = be.upload(np.array([5, 0, 7], dtype=np.uint32))
    with pytest.raises(LuminairB200Error, match="0 has no inverse"):
        be.batch_inverse(buf.ptr, buf.ptr, 3)


def test_accumulate(be):
    rng = np.random.Generator(np.random.PCG64(3))
    n = 5000
    a, b = rng.integers(0, P, size=(4, n), dtype=np.uint64), rng.integers(0, P, size=(4, n), dtype=np.uint64)
    da, db = be.upload(a.astype(np.uint32).reshape(-1)), be.upload(b.astype(np.uint32).reshape(-1))
    be.accumulate([da.at(c * n) for c in range(4)], [db.at(c * n) for c in range(4)], n)
    assert np.array_equal(be.download(da).reshape(4, n).astype(np.uint64), (a + b) % P)
