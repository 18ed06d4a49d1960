"""lb_lde_host (host buffers in, host buffers out) at 2^20 x 64 for different pipeline chunk sizes:
    python scripts/lde_host_chunks.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    from luminair_b200.backend import CudaBackend
    be = CudaBackend(0)
    n_cols, log = 64, 20
    h_in = torch.empty((n_cols, 1 << log), dtype=torch.int32, pin_memory=True)
    h_out = torch.empty((n_cols, 1 << log), dtype=torch.int32, pin_memory=True)
    a = h_in.numpy().view(np.uint32)
    a[:] = np.random.default_rng(1).integers(0, (1 << 31) - 1, size=a.shape, dtype=np.uint32)
    o = h_out.numpy().view(np.uint32)
    for chunk in (0, 1, 2, 4, 8, 16):
        for _ in range(3):
            be.lde_host(a, out=o, chunk_cols=chunk)
        assert np.array_equal(a, o)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            be.lde_host(a, out=o, chunk_cols=chunk)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"chunk_cols {chunk:2d}: min {min(ts):.3f} ms  median {sorted(ts)[5]:.3f} ms")
    be.close()


if __name__ == "__main__":
    main()
