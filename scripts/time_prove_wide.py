import os, sys, time, statistics
sys.path.insert(0, os.getcwd())
import numpy as np
from luminair_b200.backend import CudaBackend
from luminair_b200.pie import wide_graph
from luminair_b200.prover import prove, last_stage_ms
be = CudaBackend(0)
wpie = [(k, np.ascontiguousarray(v, dtype=np.uint32)) for k, v in wide_graph(20)]
dev, keep = {}, []
for name, rows in wpie:
    buf = be.upload(rows.reshape(-1)); keep.append(buf); dev[name] = (buf.ptr, rows.shape[0], rows.shape[1])
meta = [(k, None) for k, _ in wpie]
for _ in range(2): prove(meta, backend=be, device_tables=dev)
ts = []
for _ in range(6):
    t0 = time.perf_counter(); prove(meta, backend=be, device_tables=dev); ts.append((time.perf_counter() - t0) * 1e3)
print(os.environ.get("LUMINAIR_B200_LIB", "default"), "wide prove min", round(min(ts), 3), "stages", [round(x, 2) for x in last_stage_ms(be)][:3])
