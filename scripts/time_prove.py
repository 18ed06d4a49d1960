"""Time lb_prove on the BASELINE cfg-3 shape (Add 2^log rows + Inputs 2^(log+1) rows) and print the stage split."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from luminair_b200.backend import CudaBackend
from luminair_b200.prover import prove, last_stage_ms, STAGE_NAMES
from oracle import examples

log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
be = CudaBackend(0)
pie = [(n, np.ascontiguousarray(r, dtype=np.uint32)) for n, r in examples.graph_pie(log, seed=42, with_mul=False)]
dev = {}
for name, rows in pie:
    buf = be.upload(rows.reshape(-1))
    dev[name] = (buf.ptr, rows.shape[0], rows.shape[1])
    dev[name + "_buf"] = buf
for mode in ("host", "device"):
    best = None
    for r in range(reps):
        t0 = time.perf_counter()
        proof = prove(pie, backend=be, device_tables=dev if mode == "device" else None)
        dt = (time.perf_counter() - t0) * 1e3
        st = last_stage_ms(be)
        if best is None or dt < best[0]:
            best = (dt, st)
        print(f"{mode} rep {r}: {dt:.2f} ms  proof {len(proof)} B")
    print(f"== {mode}: best {best[0]:.2f} ms")
    for nm, ms in zip(STAGE_NAMES, best[1]):
        print(f"   {ms:8.2f} ms  {nm}")
