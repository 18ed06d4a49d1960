"""Pin the oracle on the one byte-level artifact the reference commits:
/root/reference/ui/demo/public/proof (copied to tests/golden/demo_proof.bin),
a bincode LuminairProof of the examples/simple graph made by an older
LuminAIR (8 component slots).  The restated *verifier* must accept it in full:
transcript, PoW, 4 Merkle decommitments, DEEP quotients, FRI, OODS check."""
import os

import pytest

from oracle import air, verifier
from oracle.channel import Blake2sChannel, grind
from oracle.fields import P, QM31
from oracle.proof import from_bincode, to_bincode


@pytest.fixture(scope="module")
def demo(golden_dir):
    data = open(os.path.join(golden_dir, "demo_proof.bin"), "rb").read()
    return data, from_bincode(data, n_components=8)


class MulEvalArtifact(air.MulEval):
    n_legacy_extra_constraints = 1


ARTIFACT_SLOTS = {0: air.AddEval, 1: MulEvalArtifact}


def test_bincode_roundtrip(demo):
    data, proof = demo
    assert len(data) == 4876
    assert to_bincode(proof) == data
    assert proof.claim[:2] == [4, 4] and proof.claim[2:] == [None] * 6
    s = proof.proof
    assert (s.config.pow_bits, s.config.log_blowup_factor, s.config.log_last_layer_degree_bound, s.config.n_queries) == (5, 1, 0, 3)
    assert [len(t) for t in s.sampled_values] == [0, 31, 24, 4]
    assert s.commitments[0].hex() == "69217a3079908094e11121d042354a7c1f55b6482ca1a51e1b250dfd1ed0eef9"


def test_verifier_accepts_reference_proof(demo):
    _, proof = demo
    ch = verifier.verify(proof, channel_variant="legacy", slot_evals=ARTIFACT_SLOTS)
    kinds = [k for k, _ in ch.log]
    # transcript of SURVEY.md Appendix C
    assert kinds == ["root", "u64", "u64", "root", "felts", "felts", "root", "root", "felts",
                     "root", "root", "root", "root", "root", "felts", "u64"]
    assert ch.log[-1] == ("u64", 2)
    assert ch.trailing_zeros() == 5


def test_current_mul_schema_is_rejected_only_at_oods(demo):
    """With today's 9-constraint Mul the composition powers shift -> OodsNotMatching;
    everything below the AIR (PCS, FRI, Merkle, PoW) still verifies."""
    _, proof = demo
    with pytest.raises(verifier.VerificationError, match="OodsNotMatching"):
        verifier.verify(proof, slot_evals={0: air.AddEval, 1: air.MulEval})
    verifier.verify(proof, slot_evals={0: air.AddEval, 1: air.MulEval}, _skip_oods=True)


def test_v2_channel_variant_rejected(demo):
    _, proof = demo
    with pytest.raises(Exception):
        verifier.verify(proof, channel_variant="v2", slot_evals=ARTIFACT_SLOTS)


def test_tampering_detected(demo):
    data, _ = demo
    for off in (200, 1500, 3000, 4800):
        bad = bytearray(data)
        bad[off] ^= 1
        with pytest.raises(Exception):
            verifier.verify(from_bincode(bytes(bad), 8), slot_evals=ARTIFACT_SLOTS)


def test_grind_and_queries(demo):
    _, proof = demo
    ch = verifier.verify(proof, slot_evals=ARTIFACT_SLOTS)
    # replay up to just before the nonce
    ch2 = Blake2sChannel("legacy")
    sp = proof.proof
    ch2.mix_root(sp.commitments[0]); ch2.mix_u64(4); ch2.mix_u64(4); ch2.mix_root(sp.commitments[1])
    for c in proof.interaction_claim[:2]:
        ch2.mix_felts([c])
    ch2.mix_root(sp.commitments[2]); ch2.mix_root(sp.commitments[3])
    ch2.mix_felts([v for t in sp.sampled_values for c in t for v in c])
    ch2.mix_root(sp.fri_proof.first_layer.commitment)
    for l in sp.fri_proof.inner_layers:
        ch2.mix_root(l.commitment)
    ch2.mix_felts(sp.fri_proof.last_layer_poly)
    assert grind(ch2, 5) == 2 == sp.proof_of_work
    ch2.mix_u64(2)
    assert ch2.digest == ch.digest
    from oracle.fri import generate_queries
    assert generate_queries(ch2, 6, 3) == [20, 25, 40]


def test_claimed_sums_from_graph_semantics(demo):
    """examples/simple (crates/../examples/simple/src/main.rs:15-22): a=[1,2,3,4],
    b=[10,20,30,40], w=-1; c=a*b (node 3), d=c+w (node 4), e=c*d (node 5); Fixed<12>.
    The Add claimed sum is sum_i -1/(c_i + 3a - z) + 1/(d_i + 4a - z) -> pins relation
    draw order, combine(), multiplicity signs and node ids."""
    _, proof = demo
    sp = proof.proof
    ch = Blake2sChannel("legacy")
    ch.mix_root(sp.commitments[0]); ch.mix_u64(4); ch.mix_u64(4); ch.mix_root(sp.commitments[1])
    node = air.RelationElements.draw(ch, 2)
    cvals = [x * y * 4096 for x, y in zip([1, 2, 3, 4], [10, 20, 30, 40])]
    dvals = [c - 4096 for c in cvals]
    s = QM31()
    for c, d in zip(cvals, dvals):
        s = s - node.combine([c, 3]).inv() + node.combine([d, 4]).inv()
    assert s == proof.interaction_claim[0]
    assert (proof.interaction_claim[0] + proof.interaction_claim[1]).is_zero()
