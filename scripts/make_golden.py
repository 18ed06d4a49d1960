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
from oracle import pie as piemod
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
    "all_components_n24.proof.bin": ("oracle.pie.all_components_graph(n=24, seed=3)  [17 components, LUTs of 2^8..2^15 rows]",
                                     lambda: piemod.all_components_graph(n=24, seed=3)),
    "mlp_2_8_8_1.proof.bin": ("oracle.pie.mlp_graph(widths=(2, 8, 8, 1))  [BASELINE cfg 4 shape at reduced width]",
                              lambda: piemod.mlp_graph(widths=(2, 8, 8, 1))),
}
for name, (desc, mk) in cases.items():
    made = mk()
    pie, pre = made if isinstance(made, tuple) else (made, ())
    data = to_bincode(prover.prove(pie, preprocessed=pre))
    open(os.path.join(G, name), "wb").write(data)
    meta[name] = {"source": "oracle/prover.py (CPU restatement), default PcsConfig, legacy channel", "pie": desc}
# fixtures at the benchmark sizes: made by the compiled CPU prover (oracle/c/cpu_prover), which tests/test_cpu_prover.py pins
# against the reference's committed proof and against the numpy oracle; the GPU tests compare lb_prove with these bytes
if "--large" in sys.argv:
    from oracle import cpu_prover
    large = {}
    big = {
        "cfg3_add_log20.proof.bin": ("oracle.pie.synthetic_add_graph_pie(20, seed=42)  [BASELINE configs[2]: Add 2^20 x 15 + Inputs 2^21 x 7]",
                                     lambda: (piemod.synthetic_add_graph_pie(20, seed=42), ())),
        "wide_log20.proof.bin": ("oracle.pie.wide_graph(20)  [headline trace shape: 2^20 rows x 61 main-trace columns + Inputs 2^21 x 7]",
                                 lambda: (piemod.wide_graph(20), ())),
        "all_components_log16.proof.bin": ("oracle.pie.all_components_graph(n=65536, seed=3)  [17 components, 2^16 .. 2^17 rows, LUTs 2^8 .. 2^15]",
                                           lambda: piemod.all_components_graph(n=1 << 16, seed=3)),
    }
    for name, (desc, mk) in big.items():
        pie, pre = mk()
        data = cpu_prover.prove(pie, preprocessed=pre)
        open(os.path.join(G, name), "wb").write(data)
        large[name] = {"source": "oracle/c/cpu_prover (compiled CPU restatement), default PcsConfig, legacy channel", "pie": desc,
                       "sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
    json.dump(large, open(os.path.join(G, "large.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(large, indent=1))
for name in meta:
    meta[name]["sha256"] = hashlib.sha256(open(os.path.join(G, name), "rb").read()).hexdigest()
    meta[name]["bytes"] = os.path.getsize(os.path.join(G, name))
json.dump(meta, open(os.path.join(G, "golden.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(meta, indent=1))
