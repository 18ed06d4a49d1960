"""Launch sequence for the ncu captures of the cfg-2 CFFT step: 3 warm round trips, then one more (4 kernels: interpolate =
low + high pass, evaluate = high + low pass).
  ncu --set full --clock-control none --import-source on -k regex:cfft_ -s 12 -c 4 -o gpurun_out/prof python scripts/profile_cfft.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from luminair_b200.backend import ColumnBatch, CudaBackend

P = (1 << 31) - 1
log, ncols = int(os.environ.get("LOG", 20)), int(os.environ.get("NCOLS", 64))
be = CudaBackend(0)
rng = np.random.Generator(np.random.PCG64(20260101))
host = rng.integers(0, P, size=(ncols, 1 << log), dtype=np.uint64).astype(np.uint32)
cb = ColumnBatch(be.upload(host.reshape(-1)), ncols, log)
be.precompute_twiddles(log)
for _ in range(4):
    be.interpolate(cb)
    be.evaluate(cb, cb)
be.sync()
assert np.array_equal(be.download(cb.buf).reshape(ncols, -1), host)
print("ok")
