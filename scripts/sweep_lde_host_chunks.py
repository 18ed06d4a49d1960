import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from luminair_b200.backend import CudaBackend
be = CudaBackend(0)
P=(1<<31)-1
n_cols, log = 64, 20
t = torch.empty((n_cols, 1<<log), dtype=torch.int32, pin_memory=True)
o = torch.empty((n_cols, 1<<log), dtype=torch.int32, pin_memory=True)
h = t.numpy().view(np.uint32); out = o.numpy().view(np.uint32)
h[:] = np.random.default_rng(1).integers(0, P, size=h.shape, dtype=np.uint64).astype(np.uint32)
import inspect
print(inspect.signature(be.lde_host))
for chunk in (0, 16, 8, 4, 2, 1):
    for _ in range(2): be.lde_host(h, out=out, chunk_cols=chunk)
    t0=time.perf_counter()
    for _ in range(5): be.lde_host(h, out=out, chunk_cols=chunk)
    print("chunk_cols", chunk, "ms", (time.perf_counter()-t0)/5*1e3, "ok", np.array_equal(h,out))
