"""Stage split of the a + b proof for every size 2^10 .. 2^22 (device-resident tables): shows where the latency-bound part of each
stage ends and the throughput-bound part begins.   python scripts/stage_by_size.py [lo hi]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from luminair_b200.backend import CudaBackend
from luminair_b200.prover import STAGE_NAMES, last_stage_ms, prove
from luminair_b200.trace import DeviceGraphTrace
from luminair_b200.workloads import build_add_graph, synthetic_add_graph_inputs

lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10, 22)
be = CudaBackend(0)
print("log | total ms | " + " | ".join(n.split(":")[0].split("(")[0].strip() for n in STAGE_NAMES))
for log in range(lo, hi + 1):
    rec = build_add_graph(DeviceGraphTrace(be), *synthetic_add_graph_inputs(log, seed=42))
    meta, dev, _ = rec.finish()
    best = None
    for _ in range(7):
        t0 = time.perf_counter()
        prove(meta, backend=be, device_tables=dev)
        dt = (time.perf_counter() - t0) * 1e3
        if best is None or dt < best[0]:
            best = (dt, last_stage_ms(be))
    print(f"{log:3d} | {best[0]:8.3f} | " + " | ".join(f"{x:6.3f}" for x in best[1]), flush=True)
    del rec, dev
be.close()
