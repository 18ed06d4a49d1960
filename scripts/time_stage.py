"""Stage split of the cfg-3 proof for the library named by LUMINAIR_B200_LIB (A/B runs of kernel variants)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from luminair_b200.backend import CudaBackend
from luminair_b200.prover import prove, last_stage_ms, STAGE_NAMES
from luminair_b200.trace import DeviceGraphTrace
from luminair_b200.workloads import build_add_graph, build_wide, synthetic_add_graph_inputs
log = int(os.environ.get("LOG", 20))
be = CudaBackend(0)
a, b = synthetic_add_graph_inputs(log, seed=42)
recs = {"cfg3": build_add_graph(DeviceGraphTrace(be), a, b), "wide": build_wide(DeviceGraphTrace(be), log)}
line = os.path.basename(os.environ.get("LUMINAIR_B200_LIB", "default"))
for name, rec in recs.items():
    meta, dev, _ = rec.finish()
    best = None
    for _ in range(6):
        t0 = time.perf_counter(); prove(meta, backend=be, device_tables=dev); dt = (time.perf_counter() - t0) * 1e3
        if best is None or dt < best[0]: best = (dt, last_stage_ms(be))
    line += f" | {name} {best[0]:.2f} ms  oods {best[1][3]:.3f} deep {best[1][4]:.3f} cq {best[1][2]:.3f} fri {best[1][5]:.3f}"
print(line)
