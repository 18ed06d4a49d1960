"""The C restatement (cpu_baseline) agrees bit-for-bit with the pinned numpy oracle."""
import numpy as np
import pytest

from oracle import cbind, cfft, merkle
from oracle.circle import CanonicCoset
from oracle.fields import P


@pytest.mark.parametrize("log", [1, 2, 5, 9, 13])
def test_c_cfft_matches_numpy(log):
    rng = np.random.Generator(np.random.PCG64(log))
    v = rng.integers(0, P, size=(3, 1 << log), dtype=np.uint64).astype(np.uint32)
    dom = CanonicCoset(log).circle_domain()
    c = cbind.CpuCfft(log)
    w = v.copy()
    c.interpolate(w)
    assert np.array_equal(w, cfft.interpolate(v, dom).astype(np.uint32))
    c.evaluate(w)
    assert np.array_equal(w, v)


def test_c_merkle_layer_matches_numpy():
    rng = np.random.Generator(np.random.PCG64(1))
    cols = [rng.integers(0, P, size=32, dtype=np.uint64).astype(np.uint32) for _ in range(19)]
    want = merkle.MerkleProver.commit(cols)
    leaf = cbind.merkle_layer(None, cols, 32)
    assert leaf.astype("<u4").tobytes() == b"".join(want.layers[5])
    l4 = cbind.merkle_layer(leaf, [], 16)
    assert l4.astype("<u4").tobytes() == b"".join(want.layers[4])
    assert cbind.merkle_layer(None, [], 1).astype("<u4").tobytes() == merkle.MerkleProver.commit([]).root()
