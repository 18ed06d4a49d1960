"""The compiled CPU prover (oracle/c/cpu_prover: packed AVX-512 / AVX2 + OpenMP restatement of the reference's CPU path) is
pinned the same way the numpy oracle is: it reproduces the reference's committed proof byte for byte, equals the numpy
oracle's bytes on every fixture / PcsConfig / channel variant, and its kernels equal the numpy ones.  It is the checker for
proofs at sizes numpy cannot reach (tests/golden/*_log20*) and the CPU baseline of bench.py."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cfft as ocfft
from oracle import cpu_prover as cp
from oracle import examples, merkle as omerkle, pie as piemod, prover as oprover
from oracle.circle import CanonicCoset
from oracle.fields import P, QM31
from oracle.proof import PcsConfig, to_bincode


def test_reproduces_the_reference_committed_proof(golden_dir):
    """ui/demo/public/proof (Add + Mul at log 4, 8-slot claim schema, legacy channel): all 4876 bytes."""
    ref = open(os.path.join(golden_dir, "demo_proof.bin"), "rb").read()
    got = cp.prove(examples.simple_pie("artifact"), n_slots=8, slot_of={"add": 0, "mul": 1}, air_era="artifact",
                   draw_lookup_elements=False)
    assert got == ref


@pytest.mark.parametrize("name", ["simple_current.proof.bin", "graph_log6_mul.proof.bin", "reduce_log5.proof.bin",
                                  "all_components_n24.proof.bin", "mlp_2_8_8_1.proof.bin"])
def test_reproduces_the_numpy_oracle_fixtures(golden_dir, name):
    from test_golden import CASES
    pie, pre = CASES[name]()
    assert cp.prove(pie, preprocessed=pre) == open(os.path.join(golden_dir, name), "rb").read()


@pytest.mark.parametrize("cfg", [PcsConfig(8, 2, 1, 5), PcsConfig(3, 1, 2, 1), PcsConfig(12, 3, 0, 40)])
@pytest.mark.parametrize("variant", ["legacy", "v2"])
def test_pcs_configs_and_channel_variants_equal_numpy(cfg, variant):
    pie = examples.graph_pie(7, seed=1, with_mul=True)
    assert cp.prove(pie, config=cfg, channel_variant=variant) == to_bincode(oprover.prove(pie, config=cfg, channel_variant=variant))


def test_mixed_heights_and_extended_domain_equal_numpy():
    """LUT tables larger than the consumers' traces (need_to_extend), tables of many heights, one thread and all threads."""
    g = piemod.GraphTrace()
    x = g.input(np.random.Generator(np.random.PCG64(11)).integers(-8, 8, 40))
    g.less_than(g.exp2(x), x)
    pie, pre = g.finish()
    want = to_bincode(oprover.prove(pie, preprocessed=pre))
    assert cp.prove(pie, preprocessed=pre, n_threads=1) == want
    assert cp.prove(pie, preprocessed=pre) == want


def test_unsatisfied_constraints_and_empty_tables():
    pie = examples.simple_pie("current")
    bad = pie[0][1].copy()
    bad[1, 11] = int(bad[1, 11]) + 1
    with pytest.raises(cp.CpuProvingError, match="ConstraintsNotSatisfied"):
        cp.prove([("add", bad)] + pie[1:])
    with pytest.raises(ValueError, match="EmptyTrace"):
        cp.prove([("add", np.zeros((0, 15), dtype=np.uint32))])


@pytest.mark.parametrize("log", [1, 3, 4, 5, 6, 9, 12, 13, 15, 18])
def test_cfft_equals_numpy(log):
    rng = np.random.Generator(np.random.PCG64(log))
    v = rng.integers(0, P, size=(3, 1 << log), dtype=np.uint64)
    a = v.astype(np.uint32)
    cp.cfft(a, forward=False)
    dom = CanonicCoset(log).circle_domain()
    want = ocfft.interpolate(v, dom).astype(np.uint32)
    assert np.array_equal(a, want)
    cp.cfft(a, forward=True)
    assert np.array_equal(a, v.astype(np.uint32))
    # LDE: coefficients zero-extended to twice the size
    ext = np.zeros((3, 2 << log), dtype=np.uint32)
    ext[:, : 1 << log] = want
    cp.cfft(ext, forward=True)
    assert np.array_equal(ext, ocfft.evaluate(want.astype(np.uint64), CanonicCoset(log + 1).circle_domain()).astype(np.uint32))


def test_merkle_and_eval_at_point_equal_numpy():
    rng = np.random.Generator(np.random.PCG64(5))
    cols = [rng.integers(0, P, size=1 << l, dtype=np.uint64).astype(np.uint32) for l in (9, 9, 7, 7, 7, 3, 10, 1, 0)]
    assert cp.merkle_root(cols) == omerkle.MerkleProver.commit(cols).root()
    wide = [rng.integers(0, P, size=1 << 6, dtype=np.uint64).astype(np.uint32) for _ in range(70)]  # > 4 message blocks
    assert cp.merkle_root(wide) == omerkle.MerkleProver.commit(wide).root()
    pt = [int(x) for x in rng.integers(1, P, size=8)]
    for log in (0, 1, 4, 9, 10, 14):
        polys = [rng.integers(0, P, size=1 << log, dtype=np.uint64).astype(np.uint32) for _ in range(3)]
        got = cp.eval_at_point(polys, pt)
        for c, g in zip(polys, got):
            want = ocfft.eval_at_point(c.astype(np.uint64), QM31(*pt[:4]), QM31(*pt[4:]))
            assert tuple(int(x) for x in g) == want.tup()


def test_large_fixtures_are_what_the_cpu_prover_produces(golden_dir):
    """The >= 2^16 fixtures (tests/golden/large.json) are regenerated here at the sizes that take seconds; the 2^20 ones are
    checked for integrity only (scripts/make_golden.py --large regenerates them in ~1 minute)."""
    meta = json.load(open(os.path.join(golden_dir, "large.json")))
    for name, m in meta.items():
        data = open(os.path.join(golden_dir, name), "rb").read()
        assert len(data) == m["bytes"] and hashlib.sha256(data).hexdigest() == m["sha256"], name
    pie, pre = piemod.all_components_graph(n=1 << 16, seed=3)
    assert cp.prove(pie, preprocessed=pre) == open(os.path.join(golden_dir, "all_components_log16.proof.bin"), "rb").read()
    # a 2^16 cross-check of the two CPU restatements themselves (numpy takes ~15 s)
    pie = piemod.synthetic_add_graph_pie(16, seed=42)
    assert cp.prove(pie) == to_bincode(oprover.prove(pie))


@pytest.mark.parametrize("seed", range(8))
def test_random_graphs_equal_numpy(seed):
    """Random operator graphs over all component kinds (the generator of tests/test_gpu_random_graphs.py): the two CPU
    restatements - numpy and the compiled prover - produce the same proof bytes."""
    from test_gpu_random_graphs import random_graph
    pie, pre = random_graph(100 + seed)
    assert cp.prove(pie, preprocessed=pre) == to_bincode(oprover.prove(pie, preprocessed=pre))
