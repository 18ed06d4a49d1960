"""Print the one-line summary of a scripts/run_sharded_prove.py JSON result."""
import json, sys
d = json.load(open(sys.argv[1]))
for k in ("cfg3_add", "wide"):
    v = d[k]
    print(d["world"], k, "eq_single", v["proof_equals_single_device"], "eq_fixture", v["proof_equals_cpu_prover_fixture"],
          "ms", round(v["ms_sharded_max_over_ranks"], 2), "vs 1 GPU", round(v["ms_single_gpu"], 2), "x", round(v["speedup"], 2),
          "MB sent", round(v["nccl"]["bytes_sent"] / 1e6, 1))
    print("   ", {a[:14]: b for a, b in v["stages_ms_rank"].items()})
