import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from luminair_b200.backend import CudaBackend
from oracle.pie import synthetic_add_graph_pie
from luminair_b200.prover import prove, last_stage_ms, STAGE_NAMES
from oracle import verifier as ov
from oracle.proof import from_bincode
be = CudaBackend(0)
for log in (22, 23):
    pie = synthetic_add_graph_pie(log, seed=1)
    prove(pie, backend=be)
    t0 = time.perf_counter(); proof = prove(pie, backend=be); dt = (time.perf_counter() - t0) * 1e3
    print(f"log {log}: {dt:.1f} ms (host tables), {len(proof)} B", flush=True)
    t0 = time.perf_counter(); ov.verify(from_bincode(proof)); print(f"   oracle verifier accepted in {time.perf_counter()-t0:.2f} s", flush=True)
