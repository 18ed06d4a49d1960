"""gen_trace on the device + prove() of the wide graph (2^log rows x 61 main-trace columns) and of the all-components graph, for
the ncu launch list / captures of the trace emitters:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_trace.csv python scripts/profile_trace.py 20
  ncu --set full --clock-control none --import-source on -k regex:trace_ -c 12 -o gpurun_out/prof_trace python scripts/profile_trace.py 20"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pie as piemod
from luminair_b200.backend import CudaBackend
from luminair_b200.prover import prove
from luminair_b200.trace import DeviceGraphTrace

log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
be = CudaBackend(0)
hg = piemod.build_all_components(piemod.GraphTrace(), 1 << max(log - 4, 4), 3)
hg.finish()
for r in range(reps + 1):  # first repetition warms the pools / twiddles
    t0 = time.perf_counter()
    dg = piemod.build_all_components(DeviceGraphTrace(be), 1 << max(log - 4, 4), 3)
    meta, dev, _ = dg.finish(hg.layouts)
    be.sync()
    t1 = time.perf_counter()
    proof = prove(meta, backend=be, device_tables=dev, preprocessed=dg.preprocessed)
    t2 = time.perf_counter()
    print(f"all components, 2^{max(log - 4, 4)} elements: gen_trace {1e3 * (t1 - t0):.2f} ms, prove {1e3 * (t2 - t1):.2f} ms, {len(proof)} B")
    t0 = time.perf_counter()
    dg = piemod.build_wide(DeviceGraphTrace(be), log)
    meta, dev, _ = dg.finish()
    be.sync()
    t1 = time.perf_counter()
    proof = prove(meta, backend=be, device_tables=dev)
    t2 = time.perf_counter()
    print(f"wide, 2^{log} rows: gen_trace (incl. host rng + upload) {1e3 * (t1 - t0):.2f} ms, prove {1e3 * (t2 - t1):.2f} ms, {len(proof)} B")
