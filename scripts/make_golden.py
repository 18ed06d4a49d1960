"""Regenerate the fixtures under tests/golden/ (run in the build container, where /root/reference is mounted).

  demo_proof.bin        byte copy of the reference's committed artifact /root/reference/ui/demo/public/proof
                        (the only known-answer vector the reference holds for this path)
  *.proof.bin           proofs produced by the CPU restatement (oracle/prover.py) for small trace tables; they pin the
                        oracle against accidental change and give the GPU tests byte targets that need no oracle run
  golden.json           sha256 of every fixture + how it was made
"""
import hashlib, json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luminair_b200 import pie as piemod
from oracle import examples, prover
from oracle.proof import to_bincode

G = os.path.join(ROOT, "tests", "golden")
meta = {}
ref = "/root/reference/ui/demo/public/proof"
if os.path.exists(ref):
    shutil.copyfile(ref, os.path.join(G, "demo_proof.bin"))
meta["demo_proof.bin"] = {"source": "copy of ui/demo/public/proof in the reference repository", "pie": "examples.simple_pie('artifact')"}
cases = {
    "simple_current.proof.bin": ("examples.simple_pie('current')", lambda: examples.simple_pie("current")),
    "graph_log6_mul.proof.bin": ("examples.graph_pie(6, seed=6, with_mul=True)", lambda: examples.graph_pie(6, seed=6, with_mul=True)),
    "reduce_log5.proof.bin": ("examples.reduce_pie(5, 2, seed=5)", lambda: examples.reduce_pie(5, 2, seed=5)),
    # graphs with lookup tables: (pie, preprocessed LUT columns)
    "all_components_n24.proof.bin": ("luminair_b200.pie.all_components_graph(n=24, seed=3)  [17 components, LUTs of 2^8..2^15 rows]",
                                     lambda: piemod.all_components_graph(n=24, seed=3)),
    "mlp_2_8_8_1.proof.bin": ("luminair_b200.pie.mlp_graph(widths=(2, 8, 8, 1))  [BASELINE cfg 4 shape at reduced width]",
                              lambda: piemod.mlp_graph(widths=(2, 8, 8, 1))),
}
for name, (desc, mk) in cases.items():
    made = mk()
    pie, pre = made if isinstance(made, tuple) else (made, ())
    data = to_bincode(prover.prove(pie, preprocessed=pre))
    open(os.path.join(G, name), "wb").write(data)
    meta[name] = {"source": "oracle/prover.py (CPU restatement), default PcsConfig, legacy channel", "pie": desc}
for name in meta:
    meta[name]["sha256"] = hashlib.sha256(open(os.path.join(G, name), "rb").read()).hexdigest()
    meta[name]["bytes"] = os.path.getsize(os.path.join(G, name))
json.dump(meta, open(os.path.join(G, "golden.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(meta, indent=1))
