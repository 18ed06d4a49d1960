"""Oracle coverage of all 17 LuminAIR components (CPU only): the restated prover's proofs pass the restated
verifier for graphs that use every component, LUT tables larger and smaller than the consuming trace
(``need_to_extend`` path of the constraint framework), and a broken witness is rejected.

Parity note: Recip/Sqrt/Rem use numerair helpers that the reference's committed proof does not cover
(``oracle/air.py``: "parity unpinned"); these tests pin self-consistency (prover <-> verifier), not reference bytes.
"""
import numpy as np
import pytest

from oracle import pie as piemod
from oracle import prover as oprover
from oracle import verifier as overifier
from oracle.proof import from_bincode, to_bincode


def _meta(pre):
    return [(cid, len(v).bit_length() - 1) for cid, v in pre]


def test_all_components_graph_shape():
    """The 17-component proof itself is pinned by tests/golden/all_components_n24.proof.bin (test_golden.py)."""
    pie, pre = piemod.all_components_graph(n=24, seed=3)
    assert [k for k, _ in pie] == ["add", "mul", "recip", "sin", "sin_lookup", "sum_reduce", "max_reduce", "sqrt", "rem",
                                   "exp2", "exp2_lookup", "log2", "log2_lookup", "less_than", "range_check_lookup",
                                   "inputs", "contiguous"]
    assert [cid for cid, _ in pre] == ["sin_lut_0", "sin_lut_1", "exp2_lut_0", "exp2_lut_1", "log2_lut_0", "log2_lut_1",
                                       "range_check_8_column_0"]


def test_lut_smaller_than_trace():
    """An Exp2 node with more rows than its table: no extension, evaluation domain = the trace's LDE."""
    g = piemod.GraphTrace()
    rng = np.random.Generator(np.random.PCG64(11))
    x = g.input(rng.integers(-8, 8, 64))  # 16 distinct raw values -> table of 16 rows, trace of 64
    g.exp2(x)
    pie, pre = g.finish()
    assert dict(pie)["exp2"].shape[0] == 64 and len(pre[0][1]) == 16
    lp = oprover.prove(pie, preprocessed=pre)
    overifier.verify(from_bincode(to_bincode(lp)), preprocessed=_meta(pre))


@pytest.mark.parametrize("kind,col", [("recip", 9), ("sqrt", 8), ("rem", 12), ("less_than", 13), ("exp2", 8)])
def test_broken_witness_is_rejected(kind, col):
    pie, pre = piemod.all_components_graph(n=16, seed=5)
    pie = [(k, v.copy()) for k, v in pie]
    tbl = dict(pie)[kind]
    tbl[1, col] = (int(tbl[1, col]) + 1) % piemod.P
    with pytest.raises((oprover.ProvingError, overifier.VerificationError)):
        lp = oprover.prove(pie, preprocessed=pre)
        # a LogUp-only break (exp2 output not in the table) passes the AIR but fails the verifier's sum check
        overifier.verify(from_bincode(to_bincode(lp)), preprocessed=_meta(pre))


def test_mlp_graph_shape():
    """BASELINE cfg 4 shape; the reduced-width proof is pinned by tests/golden/mlp_2_8_8_1.proof.bin."""
    pie, pre = piemod.mlp_graph()
    assert [k for k, _ in pie] == ["add", "mul", "recip", "sum_reduce", "exp2", "exp2_lookup", "inputs"]
    assert dict(pie)["mul"].shape[0] == 2 * 64 + 64 * 64 + 64 + 4 * 64
