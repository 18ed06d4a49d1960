"""Stage split of the cfg-4 MLP proof (2-64-64-1, tanh via the Exp2 table)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luminair_b200.backend import CudaBackend
from luminair_b200.pie import mlp_graph
from luminair_b200.prover import prove, last_stage_ms, STAGE_NAMES
be = CudaBackend(0)
pie, pre = mlp_graph()
for _ in range(3): prove(pie, backend=be, preprocessed=pre)
best = None
for _ in range(8):
    t0 = time.perf_counter(); prove(pie, backend=be, preprocessed=pre); dt = (time.perf_counter() - t0) * 1e3
    if best is None or dt < best[0]: best = (dt, last_stage_ms(be))
print(f"mlp prove best {best[0]:.2f} ms")
for nm, ms in zip(STAGE_NAMES, best[1]): print(f"   {ms:7.2f} ms  {nm}")
