"""CircuitSettings (crates/air/src/settings.rs): bincode round trip, the LUT columns derived from it, and the one
settings file the reference commits (ui/demo/public/settings: a single 0x00 byte, the one-Option schema of that revision)."""
import numpy as np
import pytest

from oracle import pie as piemod
from luminair_b200.settings import CircuitSettings, Lookup, RangeCheckLookup


def test_empty_settings_are_four_none_tags():
    assert CircuitSettings().to_bincode() == b"\x00\x00\x00\x00"
    assert CircuitSettings.from_bincode(b"\x00\x00\x00\x00").preprocessed_columns() == []


def test_roundtrip_and_columns_match_graph_trace():
    pie, pre = piemod.all_components_graph(n=24, seed=3)
    s = CircuitSettings.from_graph_trace(pie, pre)
    cols = s.preprocessed_columns()
    assert [c for c, _ in cols] == [c for c, _ in pre]
    for (cid, v), (_, w) in zip(cols, pre):
        assert np.array_equal(v, w), cid
    data = s.to_bincode()
    back = CircuitSettings.from_bincode(data)
    assert back.to_bincode() == data
    assert back.sin.ranges == s.sin.ranges and back.range_check.n_bits == 8
    assert np.array_equal(back.exp2.multiplicities, np.asarray(dict(pie)["exp2_lookup"]).reshape(-1))


def test_layout_bytes():
    """Field order and widths: Option tag, Vec<(i64, i64)>, u32 log_size, Vec<u32>."""
    s = CircuitSettings(exp2=Lookup([(-2, 13)], 4, np.arange(16, dtype=np.uint32)))
    b = s.to_bincode()
    assert b[0] == 0 and b[1] == 1                       # sin None, exp2 Some
    assert b[2:10] == (1).to_bytes(8, "little")          # one range
    assert b[10:18] == (-2).to_bytes(8, "little", signed=True) and b[18:26] == (13).to_bytes(8, "little")
    assert b[26:30] == (4).to_bytes(4, "little") and b[30:38] == (16).to_bytes(8, "little")
    assert len(b) == 38 + 64 + 2                         # multiplicities, then log2 None, range_check None
    rc = CircuitSettings(range_check=RangeCheckLookup(8, 8, np.zeros(256, dtype=np.uint32))).to_bincode()
    assert rc[:3] == b"\x00\x00\x00" and rc[3] == 1 and rc[4:8] == (8).to_bytes(4, "little")


def test_bad_bytes_are_rejected():
    with pytest.raises(ValueError):
        CircuitSettings.from_bincode(b"\x00\x00\x00\x00\x00")
    with pytest.raises(ValueError):
        CircuitSettings.from_bincode(b"\x02\x00\x00\x00")


@pytest.mark.gpu
def test_prove_with_settings_equals_prove_with_columns():
    from luminair_b200.prover import prove
    pie, pre = piemod.mlp_graph(widths=(2, 8, 8, 1))
    s = CircuitSettings.from_bincode(CircuitSettings.from_graph_trace(pie, pre).to_bincode())
    assert prove(pie, settings=s) == prove(pie, preprocessed=pre)


def test_mlp_exp2_range_is_the_calibrated_one():
    """workloads.MLP_EXP2_RANGE (the circuit setting bench.py proves the cfg-4 MLP with) is the covering range a calibration
    run records, and the product's LUT columns equal the checker's own (both through host libm)."""
    from luminair_b200 import lookups, workloads
    g = piemod.build_mlp(piemod.GraphTrace())
    _, pre = g.finish()
    assert g.layouts["exp2"].ranges == [workloads.MLP_EXP2_RANGE]
    mine = lookups.lut_columns("exp2", lookups.LookupLayout([workloads.MLP_EXP2_RANGE]))
    for (cid, want), (cid2, got) in zip(pre, mine):
        assert cid == cid2 and np.array_equal(want, got)
    for name in ("sin", "log2"):
        lay = lookups.LookupLayout([(1, 5000)])
        olay = piemod.LookupLayout([(1, 5000)])
        for (_, a), (_, b) in zip(lookups.lut_columns(name, lay), piemod.lut_columns(name, olay)):
            assert np.array_equal(a, b)
    assert int(lookups.to_fixed(2.0 ** -13)) == 1 and int(lookups.to_fixed(-(2.0 ** -13))) == -1  # ties away from zero
