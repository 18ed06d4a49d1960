"""N > 1 path on CPU: the column-sharded commit (luminair_b200/sharded.py) with world_size 2 and 4 over gloo.
The collectives, the shard arithmetic and the top-of-tree hashing are the product's; the per-rank kernels are
replaced by the oracle (no GPU here).  The root must equal the single-process oracle commitment of ALL columns."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cfft as ocfft, merkle as omerkle
from oracle.circle import CanonicCoset
from oracle.fields import P


class OracleShardOps:
    def lde(self, trace, log_size, log_blowup):
        vals = trace.numpy().view(np.uint32).astype(np.uint64)
        coeffs = ocfft.interpolate(vals, CanonicCoset(log_size).circle_domain())
        ev = ocfft.evaluate(coeffs, CanonicCoset(log_size + log_blowup).circle_domain())
        return torch.from_numpy(ev.astype(np.uint32).view(np.int32).copy())

    def merkle_layer(self, log_size, prev, cols):
        n = 1 << log_size
        out = np.empty((n, 8), dtype="<u4")
        pv = None if prev is None else prev.numpy().view(np.uint32)
        cv = None if cols is None else cols.numpy().view(np.uint32)
        for i in range(n):
            children = None if pv is None else (pv[2 * i].astype("<u4").tobytes(), pv[2 * i + 1].astype("<u4").tobytes())
            vals = [] if cv is None else cv[:, i]
            out[i] = np.frombuffer(omerkle.hash_node(children, vals), dtype="<u4")
        return torch.from_numpy(out.view(np.int32).copy())

    def sync(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _trace(n_cols, log, seed=5):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, P, size=(n_cols, 1 << log), dtype=np.uint64).astype(np.uint32)


def _worker(rank, world, port, n_cols, log, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from luminair_b200.sharded import column_range, sharded_commit
        full = _trace(n_cols, log)
        lo, hi = column_range(n_cols, rank, world)
        local = torch.from_numpy(full[lo:hi].view(np.int32).copy())
        timings = {}
        root = sharded_commit(OracleShardOps(), local, log, 1, timings=timings)
        q.put((rank, root, timings["all_to_all_bytes_per_rank"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_commit_root_equals_single_tree(world):
    n_cols, log = 8, 5
    full = _trace(n_cols, log).astype(np.uint64)
    coeffs = ocfft.interpolate(full, CanonicCoset(log).circle_domain())
    lde = ocfft.evaluate(coeffs, CanonicCoset(log + 1).circle_domain())
    want = omerkle.MerkleProver.commit(list(lde)).root()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cols, log, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, root, a2a in got:
        assert root == want, f"rank {rank}: sharded root differs from the single-tree root"
        assert a2a == (n_cols // world) * (2 << log) * 4 * (world - 1) // world


def test_single_rank_degenerates_to_plain_commit():
    from luminair_b200.sharded import sharded_commit
    n_cols, log = 3, 4
    full = _trace(n_cols, log, seed=9)
    lde = ocfft.evaluate(ocfft.interpolate(full.astype(np.uint64), CanonicCoset(log).circle_domain()), CanonicCoset(log + 1).circle_domain())
    want = omerkle.MerkleProver.commit(list(lde)).root()
    got = sharded_commit(OracleShardOps(), torch.from_numpy(full.view(np.int32).copy()), log, 1)
    assert got == want


def test_column_range_rejects_ragged_shards():
    from luminair_b200.sharded import column_range
    assert column_range(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        column_range(10, 0, 4)
