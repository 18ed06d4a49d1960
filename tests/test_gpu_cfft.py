"""GPU parity: CFFT + Merkle through the C ABI vs the oracle (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cfft as ocfft
from oracle import merkle as omerkle
from oracle.circle import CanonicCoset, Coset
from oracle.fields import P


@pytest.fixture(scope="module")
def be():
    from luminair_b200.backend import CudaBackend
    b = CudaBackend(0)
    yield b
    b.close()


def _batch(be, arr, log):
    from luminair_b200.backend import ColumnBatch
    buf = be.upload(arr.reshape(-1))
    return ColumnBatch(buf, arr.shape[0], log)


@pytest.mark.parametrize("log", [1, 2, 3, 4, 5, 7, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22])
def test_interpolate_evaluate_vs_oracle(be, log):
    rng = np.random.Generator(np.random.PCG64(log))
    ncols = 5 if log <= 12 else (3 if log <= 18 else 2)
    vals = rng.integers(0, P, size=(ncols, 1 << log), dtype=np.uint64).astype(np.uint32)
    dom = CanonicCoset(log).circle_domain()
    cb = _batch(be, vals, log)
    be.interpolate(cb)
    got = be.download(cb.buf).reshape(ncols, -1)
    want = ocfft.interpolate(vals, dom).astype(np.uint32)
    assert np.array_equal(got, want)
    # evaluate back (same size) -> original
    from luminair_b200.backend import ColumnBatch
    out = ColumnBatch(be.alloc(ncols << log), ncols, log)
    be.evaluate(cb, out)
    assert np.array_equal(be.download(out.buf).reshape(ncols, -1), vals)
    # LDE to log+1
    out2 = ColumnBatch(be.alloc(ncols << (log + 1)), ncols, log + 1)
    be.evaluate(cb, out2)
    want2 = ocfft.evaluate(want, CanonicCoset(log + 1).circle_domain()).astype(np.uint32)
    assert np.array_equal(be.download(out2.buf).reshape(ncols, -1), want2)


def test_edge_values(be):
    log = 6
    vals = np.zeros((4, 1 << log), dtype=np.uint32)
    vals[1, :] = P - 1
    vals[2, 0] = 1
    vals[3, ::2] = P - 1
    cb = _batch(be, vals, log)
    be.interpolate(cb)
    want = ocfft.interpolate(vals, CanonicCoset(log).circle_domain()).astype(np.uint32)
    assert np.array_equal(be.download(cb.buf).reshape(4, -1), want)


def test_twiddle_tree_matches_stwo_layout(be):
    for root_log in (1, 2, 5, 9):
        got = be.export_twiddles(root_log)
        want = ocfft.precompute_twiddles(Coset.half_odds(root_log)).astype(np.uint32)
        assert np.array_equal(got, want)


def test_roundtrip_full_size_property(be):
    """BASELINE cfg 2 shape at reduced column count: 2^20 rows; interpolate∘evaluate = id."""
    from luminair_b200.backend import ColumnBatch
    log, ncols = 20, 4
    rng = np.random.Generator(np.random.PCG64(20260101))
    vals = rng.integers(0, P, size=(ncols, 1 << log), dtype=np.uint64).astype(np.uint32)
    cb = _batch(be, vals, log)
    be.interpolate(cb)
    coeffs = be.download(cb.buf).reshape(ncols, -1)
    assert coeffs.max() < P
    # linearity spot check against the oracle on one column
    want = ocfft.interpolate(vals[:1], CanonicCoset(log).circle_domain()).astype(np.uint32)
    assert np.array_equal(coeffs[:1], want)
    out = ColumnBatch(be.alloc(ncols << log), ncols, log)
    be.evaluate(cb, out)
    assert np.array_equal(be.download(out.buf).reshape(ncols, -1), vals)


def _commit_tree(be, cols_by_log):
    """cols_by_log: {log: np.ndarray[n_cols, 2^log]} -> (root bytes, layers)"""
    max_log = max(cols_by_log)
    bufs = {l: be.upload(a.reshape(-1)) for l, a in cols_by_log.items()}
    prev = None
    layers = {}
    for log in range(max_log, -1, -1):
        out = be.alloc(8 << log)
        ptrs = []
        if log in cols_by_log:
            a = cols_by_log[log]
            ptrs = [bufs[log].at(c << log) for c in range(a.shape[0])]
        be.merkle_commit_layer(log, prev.ptr if prev is not None else None, ptrs, out.ptr)
        layers[log] = out
        prev = out
    root = be.download(layers[0]).astype("<u4").tobytes()
    return root, layers


@pytest.mark.parametrize("shape", [{5: 3}, {6: 31}, {4: 16}, {7: 17, 5: 4, 2: 1}, {3: 0, 2: 2}, {19: 17, 18: 33}, {18: 4}])
def test_merkle_vs_oracle(be, shape):
    rng = np.random.Generator(np.random.PCG64(7))
    cols_by_log = {l: rng.integers(0, P, size=(n, 1 << l), dtype=np.uint64).astype(np.uint32) for l, n in shape.items()}
    cols_by_log = {l: a for l, a in cols_by_log.items()}
    flat = [a[c] for l, a in cols_by_log.items() for c in range(a.shape[0])]
    if not any(len(c) == 1 << max(shape) for c in flat):
        pytest.skip("largest layer must hold a column")
    want = omerkle.MerkleProver.commit(flat)
    root, layers = _commit_tree(be, {l: a for l, a in cols_by_log.items() if a.shape[0]})
    assert root == want.root()
    for log, buf in layers.items():
        got = be.download(buf).astype("<u4").tobytes()
        assert got == b"".join(want.layers[log])


def test_merkle_empty_tree(be):
    out = be.alloc(8)
    be.merkle_commit_layer(0, None, [], out.ptr)
    assert be.download(out).astype("<u4").tobytes().hex() == "69217a3079908094e11121d042354a7c1f55b6482ca1a51e1b250dfd1ed0eef9"


def test_gather_rows(be):
    rng = np.random.Generator(np.random.PCG64(3))
    a = rng.integers(0, P, size=(5, 64), dtype=np.uint64).astype(np.uint32)
    buf = be.upload(a.reshape(-1))
    got = be.gather_rows([buf.at(c * 64) for c in range(5)], [3, 17, 63])
    assert np.array_equal(got, a[:, [3, 17, 63]].T)


def test_evaluate_in_place(be):
    from luminair_b200.backend import ColumnBatch
    log, ncols = 14, 3
    rng = np.random.Generator(np.random.PCG64(99))
    vals = rng.integers(0, P, size=(ncols, 1 << log), dtype=np.uint64).astype(np.uint32)
    cb = _batch(be, vals, log)
    be.interpolate(cb)
    be.evaluate(cb, cb)
    assert np.array_equal(be.download(cb.buf).reshape(ncols, -1), vals)


@pytest.mark.gpu
@pytest.mark.parametrize("log_in,log_out,n_cols,chunk", [(10, 10, 7, 2), (12, 13, 5, 0), (14, 15, 9, 4)])
def test_lde_host_pipeline(log_in, log_out, n_cols, chunk):
    """lb_lde_host (chunked upload / interpolate / evaluate / download over three streams) == oracle, incl. ragged last chunk."""
    from luminair_b200.backend import CudaBackend
    from oracle import cfft as ocfft
    from oracle.circle import CanonicCoset
    from oracle.fields import P
    be = CudaBackend(0)
    rng = np.random.Generator(np.random.PCG64(log_in * 100 + n_cols))
    vals = rng.integers(0, P, size=(n_cols, 1 << log_in), dtype=np.uint64)
    coeffs = np.empty((n_cols, 1 << log_in), dtype=np.uint32)
    got = be.lde_host(vals.astype(np.uint32), log_out, coeffs_out=coeffs, chunk_cols=chunk)
    want_c = ocfft.interpolate(vals, CanonicCoset(log_in).circle_domain())
    want = ocfft.evaluate(want_c, CanonicCoset(log_out).circle_domain())
    assert np.array_equal(coeffs, want_c.astype(np.uint32))
    assert np.array_equal(got, want.astype(np.uint32))
    be.close()
