"""Largest tables the prover accepts (Inputs table of 2^24 rows at log 23): lb_prove against the compiled CPU prover, byte for
byte, and through the numpy verifier.  One-off check, too slow for the test suite:  python scripts/prove_large.py [logs...]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from luminair_b200.backend import CudaBackend
from luminair_b200.prover import prove
from oracle import cpu_prover as cp
from oracle import verifier as ov
from oracle.pie import synthetic_add_graph_pie
from oracle.proof import from_bincode

be = CudaBackend(0)
for log in [int(a) for a in sys.argv[1:]] or (22, 23):
    pie = synthetic_add_graph_pie(log, seed=1)
    prove(pie, backend=be)
    t0 = time.perf_counter()
    proof = prove(pie, backend=be)
    dt = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    want = cp.prove(pie)
    dc = (time.perf_counter() - t0) * 1e3
    print(f"log {log}: GPU {dt:.1f} ms (host tables), CPU prover {dc:.0f} ms on {cp.host_cores()} cores, {len(proof)} B, "
          f"bytes equal: {proof == want}", flush=True)
    assert proof == want
    t0 = time.perf_counter()
    ov.verify(from_bincode(proof))
    print(f"   numpy verifier accepted in {time.perf_counter() - t0:.2f} s", flush=True)
be.close()
