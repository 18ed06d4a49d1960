"""LuminairProof / CircuitSettings wire formats of the product package (luminair_b200.proof, luminair_b200.settings):
bincode <-> structure <-> JSON round trips on every committed proof, including the reference's own artifact."""
import json
import os

import numpy as np
import pytest

from luminair_b200 import proof as wire
from luminair_b200.settings import CircuitSettings, Lookup, RangeCheckLookup

FIXTURES = ["simple_current.proof.bin", "graph_log6_mul.proof.bin", "reduce_log5.proof.bin", "all_components_n24.proof.bin",
            "mlp_2_8_8_1.proof.bin", "cfg3_add_log20.proof.bin", "wide_log20.proof.bin", "all_components_log16.proof.bin"]


@pytest.mark.parametrize("name", FIXTURES)
def test_bincode_json_round_trip(golden_dir, name):
    data = open(os.path.join(golden_dir, name), "rb").read()
    p = wire.from_bincode(data)
    assert wire.to_bincode(p) == data
    text = wire.to_json(p)
    assert wire.json_to_bincode(text) == data
    doc = json.loads(text)
    assert list(doc) == ["claim", "interaction_claim", "proof"]  # LuminairProof field order, crates/prover/src/lib.rs:15-20
    assert list(doc["claim"]) == wire.COMPONENTS and list(doc["interaction_claim"]) == wire.COMPONENTS
    live = [n for n in wire.COMPONENTS if doc["claim"][n] is not None]
    assert live and all(set(doc["claim"][n]) == {"log_size", "_marker"} for n in live)
    assert all(len(doc["interaction_claim"][n]["claimed_sum"]) == 2 for n in live)
    assert len(doc["proof"]["commitments"]) == 4 and all(len(h) == 32 for h in doc["proof"]["commitments"])


def test_reference_artifact_round_trips(golden_dir):
    """ui/demo/public/proof: the 8-slot schema of the revision that made it (add, mul first)."""
    data = open(os.path.join(golden_dir, "demo_proof.bin"), "rb").read()
    p = wire.from_bincode(data, n_components=8)
    assert wire.to_bincode(p) == data
    names = wire.COMPONENTS[:2] + [f"slot{k}" for k in range(2, 8)]
    assert wire.to_bincode(wire.from_json(wire.to_json(p, names), names)) == data
    assert p.claim[:2] == [4, 4] and p.proof_of_work == 2 and p.n_queries == 3


def test_json_file_helpers(tmp_path, golden_dir):
    data = open(os.path.join(golden_dir, "simple_current.proof.bin"), "rb").read()
    path = tmp_path / "proof.json"
    wire.write_json_file(str(path), data)
    assert wire.read_json_file(str(path)) == data
    with pytest.raises(ValueError):
        wire.from_bincode(data + b"\0")
    with pytest.raises(ValueError):
        wire.from_bincode(data[:-3])


def test_circuit_settings_json_round_trip():
    rng = np.random.Generator(np.random.PCG64(3))
    s = CircuitSettings(sin=Lookup([(-5, 20), (100, 130)], 6, rng.integers(0, 9, 64).astype(np.uint32)),
                        log2=Lookup([(1, 4096)], 12, rng.integers(0, 3, 4096).astype(np.uint32)),
                        range_check=RangeCheckLookup(8, 8, rng.integers(0, 1000, 256).astype(np.uint32)))
    back = CircuitSettings.from_json(s.to_json())
    assert back.to_bincode() == s.to_bincode()
    assert CircuitSettings.from_bincode(back.to_bincode()).to_json() == s.to_json()
    doc = json.loads(s.to_json())
    assert list(doc) == ["lookups"] and list(doc["lookups"]) == ["sin", "exp2", "log2", "range_check"]
    assert doc["lookups"]["exp2"] is None and doc["lookups"]["sin"]["layout"]["ranges"][0] == [-5, 20]
    assert CircuitSettings.from_json(CircuitSettings().to_json()).to_bincode() == b"\0\0\0\0"
