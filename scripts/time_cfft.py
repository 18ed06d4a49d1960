"""Time the cfg-2 CFFT round trip (64 x 2^20) for the library named by LUMINAIR_B200_LIB; checks the round trip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from luminair_b200.backend import ColumnBatch, CudaBackend

P = (1 << 31) - 1
log, ncols = int(os.environ.get("LOG", 20)), int(os.environ.get("NCOLS", 64))
be = CudaBackend(0)
rng = np.random.Generator(np.random.PCG64(20260101))
host = rng.integers(0, P, size=(ncols, 1 << log), dtype=np.uint64).astype(np.uint32)
buf = be.upload(host.reshape(-1))
cb = ColumnBatch(buf, ncols, log)
be.precompute_twiddles(log + 1)
be.interpolate(cb); be.evaluate(cb, cb)
ok = np.array_equal(be.download(buf).reshape(ncols, -1), host)
lde = ColumnBatch(be.alloc(ncols << (log + 1)), ncols, log + 1)
for _ in range(5):
    be.interpolate(cb); be.evaluate(cb, cb)
def t(fn, reps=20):
    be.timer_start()
    for _ in range(reps): fn()
    return be.timer_stop_ms() / reps
ti = t(lambda: be.interpolate(cb)); te = t(lambda: be.evaluate(cb, cb))
be.interpolate(cb)
tl = t(lambda: be.evaluate(cb, lde))
be.evaluate(cb, cb)
rt = t(lambda: (be.interpolate(cb), be.evaluate(cb, cb)))
print(f"{os.environ.get('LUMINAIR_B200_LIB','default'):>28s} ok={ok} interp {ti:.4f} eval {te:.4f} lde {tl:.4f} roundtrip {rt:.4f} ms  frac {1073.741824/rt/6544:.3f}")
