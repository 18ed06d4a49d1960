"""The oracle *prover* reproduces the reference's committed proof byte-for-byte, and
its proofs for the current schema verify."""
import os

import numpy as np
import pytest

from oracle import air, examples, prover, verifier
from oracle.proof import from_bincode, to_bincode


class MulEvalArtifact(air.MulEval):
    n_legacy_extra_constraints = 1


ARTIFACT_SLOTS = {0: air.AddEval, 1: MulEvalArtifact}


def test_reproduces_reference_proof_bytes(golden_dir):
    ref = open(os.path.join(golden_dir, "demo_proof.bin"), "rb").read()
    lp = prover.prove(examples.simple_pie("artifact"), n_slots=8, slot_evals=ARTIFACT_SLOTS,
                      slot_of={"add": 0, "mul": 1}, channel_variant="legacy")
    out = to_bincode(lp)
    assert out == ref  # all 4876 bytes


def test_simple_current_schema_proves_and_verifies():
    lp = prover.prove(examples.simple_pie("current"))
    assert lp.claim[0] == 4 and lp.claim[1] == 4 and lp.claim[15] == 4
    assert verifier.log_sum_valid(lp.interaction_claim)
    data = to_bincode(lp)
    verifier.verify(from_bincode(data))


def test_unsatisfied_constraints_are_caught():
    pie = examples.simple_pie("current")
    bad = pie[0][1].copy()
    bad[1, 11] = (int(bad[1, 11]) + 1)  # out != lhs + rhs
    with pytest.raises(prover.ProvingError):
        prover.prove([("add", bad)] + pie[1:])


def test_empty_table_is_an_error():
    with pytest.raises(ValueError, match="EmptyTrace"):
        prover.prove([("add", np.zeros((0, 15), dtype=np.uint64))])


def test_synthetic_mul_table_log8():
    t = examples.synthetic_mul_table(8, seed=42)
    lp = prover.prove([("mul", t)])
    verifier.verify(lp)
    assert lp.claim[1] == 8
