"""Committed fixtures (tests/golden, made by scripts/make_golden.py): integrity, the oracle against them (CPU),
the CUDA prover against them (GPU) without running the oracle."""
import hashlib
import json
import os

import pytest

from oracle import pie as piemod
from oracle import examples

# name -> (pie, preprocessed LUT columns)
CASES = {
    "simple_current.proof.bin": lambda: (examples.simple_pie("current"), ()),
    "graph_log6_mul.proof.bin": lambda: (examples.graph_pie(6, seed=6, with_mul=True), ()),
    "reduce_log5.proof.bin": lambda: (examples.reduce_pie(5, 2, seed=5), ()),
    "all_components_n24.proof.bin": lambda: piemod.all_components_graph(n=24, seed=3),
    "mlp_2_8_8_1.proof.bin": lambda: piemod.mlp_graph(widths=(2, 8, 8, 1)),
}


def _meta(golden_dir):
    return json.load(open(os.path.join(golden_dir, "golden.json")))


def test_fixture_integrity(golden_dir):
    meta = _meta(golden_dir)
    assert set(CASES) | {"demo_proof.bin"} == set(meta)
    for name, m in meta.items():
        data = open(os.path.join(golden_dir, name), "rb").read()
        assert len(data) == m["bytes"] and hashlib.sha256(data).hexdigest() == m["sha256"], name


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_fixture(golden_dir, name):
    from oracle import prover, verifier
    from oracle.proof import from_bincode, to_bincode
    want = open(os.path.join(golden_dir, name), "rb").read()
    pie, pre = CASES[name]()
    assert to_bincode(prover.prove(pie, preprocessed=pre)) == want
    verifier.verify(from_bincode(want), preprocessed=[(cid, len(v).bit_length() - 1) for cid, v in pre])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_prover_reproduces_fixture(golden_dir, name):
    from luminair_b200.prover import prove
    want = open(os.path.join(golden_dir, name), "rb").read()
    pie, pre = CASES[name]()
    assert prove(pie, preprocessed=pre) == want


# fixtures at the benchmark sizes (tests/golden/large.json), made by the compiled CPU prover (scripts/make_golden.py --large;
# tests/test_cpu_prover.py pins that prover): whole-proof byte parity of lb_prove on BASELINE configs[2] at 2^20, on the
# headline 2^20 x 61-column trace shape and on a 2^16-element graph with all 17 components - no oracle run on the GPU box
LARGE = {
    "cfg3_add_log20.proof.bin": lambda: (piemod.synthetic_add_graph_pie(20, seed=42), ()),
    "wide_log20.proof.bin": lambda: (piemod.wide_graph(20), ()),
    "all_components_log16.proof.bin": lambda: piemod.all_components_graph(n=1 << 16, seed=3),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LARGE))
def test_cuda_prover_reproduces_large_fixture(golden_dir, name):
    from luminair_b200.prover import prove
    want = open(os.path.join(golden_dir, name), "rb").read()
    pie, pre = LARGE[name]()
    got = prove(pie, preprocessed=pre)
    assert len(got) == len(want)
    assert got == want
