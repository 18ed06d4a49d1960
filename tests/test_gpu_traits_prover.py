"""A proof assembled from the backend-TRAIT entry points of the C ABI (luminair_b200.traits_prover: host-side channel +
lb_interpolate_batch / lb_evaluate_batch / lb_merkle_commit_layer / lb_logup_* / lb_constraint_quotients_lut / lb_eval_at_point /
lb_accumulate_quotients / lb_fold_* / lb_grind, one call per stwo trait method) equals, byte for byte, the committed fixtures and
what lb_prove returns - the evidence for INTEGRATION.md path B (SURVEY 8b "Reality check")."""
import os

import pytest

from oracle import examples, pie as piemod

CASES = {
    "simple_current.proof.bin": lambda: (examples.simple_pie("current"), ()),
    "graph_log6_mul.proof.bin": lambda: (examples.graph_pie(6, seed=6, with_mul=True), ()),
    "reduce_log5.proof.bin": lambda: (examples.reduce_pie(5, 2, seed=5), ()),
    "all_components_n24.proof.bin": lambda: piemod.all_components_graph(n=24, seed=3),  # LUTs larger than their consumers
    "mlp_2_8_8_1.proof.bin": lambda: piemod.mlp_graph(widths=(2, 8, 8, 1)),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_trait_level_proof_equals_fixture_and_lb_prove(golden_dir, name):
    from luminair_b200.backend import CudaBackend
    from luminair_b200.prover import prove
    from luminair_b200.traits_prover import prove_with_traits
    pie, pre = CASES[name]()
    be = CudaBackend(0)
    try:
        got = prove_with_traits(pie, be, preprocessed=pre)
        assert got == open(os.path.join(golden_dir, name), "rb").read()
        assert got == prove(pie, backend=be, preprocessed=pre)
    finally:
        be.close()


@pytest.mark.gpu
def test_trait_level_proof_other_config_and_channel():
    from luminair_b200.backend import CudaBackend
    from luminair_b200.prover import PcsConfig, prove
    from luminair_b200.traits_prover import prove_with_traits
    pie = examples.graph_pie(9, seed=4, with_mul=True)
    be = CudaBackend(0)
    try:
        for cfg, variant in ((PcsConfig(8, 2, 1, 5), "legacy"), (PcsConfig(3, 1, 2, 7), "v2")):
            assert prove_with_traits(pie, be, config=cfg, channel_variant=variant) == prove(pie, backend=be, config=cfg, channel_variant=variant)
    finally:
        be.close()


def test_host_transcript_math_matches_the_checker():
    """hostmath (the twin's scalar side) against the numpy oracle: QM31 arithmetic, circle points, both channel variants."""
    import numpy as np
    from luminair_b200 import hostmath as hm
    from oracle import channel as ochannel, circle as ocircle, fields as of
    rng = np.random.Generator(np.random.PCG64(9))
    for _ in range(20):
        a, b = [tuple(int(x) for x in rng.integers(0, of.P, 4)) for _ in range(2)]
        qa, qb, oa, ob = hm.QM31(*a), hm.QM31(*b), of.QM31(*a), of.QM31(*b)
        assert (qa * qb).c == (oa * ob).tup() and (qa + qb).c == (oa + ob).tup() and (qa - qb).c == (oa - ob).tup()
        assert qa.inv().c == oa.inv().tup() and (qa * 12345).c == (oa * 12345).tup()
    for idx in (0, 1, 5, 1 << 30, (1 << 31) - 3, -7):
        assert hm.index_to_point(idx) == ocircle.index_to_point(idx)
    for variant in ("legacy", "v2"):
        c1, c2 = hm.Blake2sChannel(variant), ochannel.Blake2sChannel(variant)
        for ch in (c1, c2):
            ch.mix_root(bytes(range(32)))
            ch.mix_u64(20)
        f1, f2 = c1.draw_secure_felts(3), c2.draw_secure_felts(3)
        assert [f.c for f in f1] == [f.tup() for f in f2]
        c1.mix_felts(f1)
        c2.mix_felts(f2)
        assert c1.digest == c2.digest and c1.draw_random_bytes() == c2.draw_random_bytes()
