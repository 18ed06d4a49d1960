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

from oracle import cfft as ocfft, merkle as omerkle, quotients as oquot
from oracle.circle import CanonicCoset
from oracle.fields import P, QM31


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

    def eval_at_point(self, coeffs, log_size, point):
        px, py = QM31(*point[:4]), QM31(*point[4:])
        out = [ocfft.eval_at_point(c.astype(np.uint64), px, py).tup() for c in coeffs.numpy().view(np.uint32)]
        return torch.from_numpy(np.array(out, dtype=np.uint32).view(np.int32).copy())

    def quotients_partial(self, lde, lde_log, point, values, random_coeff, col_offset, n_cols_global):
        """Stand-in for lb_accumulate_quotients_shard: the oracle's full accumulation with every column this rank does
        not own replaced by the zero column sampled to zero (its line coefficients vanish, so it contributes nothing)."""
        pt = (QM31(*point[:4]), QM31(*point[4:]))
        n = 1 << lde_log
        mine = lde.numpy().view(np.uint32).astype(np.uint64)
        vals = values.numpy().view(np.uint32)
        cols, samples = [], []
        for j in range(n_cols_global):
            k = j - col_offset
            own = 0 <= k < mine.shape[0]
            cols.append(mine[k] if own else np.zeros(n, dtype=np.uint64))
            samples.append([(pt, QM31(*[int(x) for x in vals[k]]) if own else QM31())])
        q = oquot.accumulate_quotients(lde_log, cols, samples, QM31(*random_coeff))
        return torch.from_numpy(np.stack([np.asarray(c, dtype=np.uint32) for c in q.c]).view(np.int32).copy())

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


# ---- OODS sampling + DEEP quotient accumulation over column shards -------------------------------------
POINT = [11, 22, 33, 44, 55, 66, 77, 88]
RC = [5, 6, 7, 8]


def _quot_worker(rank, world, port, n_cols, log, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from luminair_b200.sharded import column_range, sharded_quotient_accumulation
        full = _trace(n_cols, log, seed=8).astype(np.uint64)
        lo, hi = column_range(n_cols, rank, world)
        coeffs = ocfft.interpolate(full[lo:hi], CanonicCoset(log).circle_domain())
        lde = ocfft.evaluate(coeffs, CanonicCoset(log + 1).circle_domain())
        to_t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.uint32).view(np.int32).copy())
        timings = {}
        sampled, quot = sharded_quotient_accumulation(OracleShardOps(), to_t(coeffs), to_t(lde), log, 1, POINT, RC, timings=timings)
        q.put((rank, sampled.numpy().view(np.uint32).copy(), quot.numpy().view(np.uint32).copy(), timings))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_quotients_equal_single_device(world):
    n_cols, log = 8, 4
    full = _trace(n_cols, log, seed=8).astype(np.uint64)
    coeffs = ocfft.interpolate(full, CanonicCoset(log).circle_domain())
    lde = ocfft.evaluate(coeffs, CanonicCoset(log + 1).circle_domain())
    pt = (QM31(*POINT[:4]), QM31(*POINT[4:]))
    want_s = [ocfft.eval_at_point(c, pt[0], pt[1]) for c in coeffs]
    want_q = oquot.accumulate_quotients(log + 1, list(lde), [[(pt, v)] for v in want_s], QM31(*RC))
    want_q = np.stack([np.asarray(c, dtype=np.uint32) for c in want_q.c])
    if world == 1:
        from luminair_b200.sharded import sharded_quotient_accumulation
        to_t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.uint32).view(np.int32).copy())
        sampled, quot = sharded_quotient_accumulation(OracleShardOps(), to_t(coeffs), to_t(lde), log, 1, POINT, RC)
        got = [(0, sampled.numpy().view(np.uint32), quot.numpy().view(np.uint32), {})]
    else:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_quot_worker, args=(r, world, port, n_cols, log, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = [q.get(timeout=120) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    for rank, sampled, quot, timings in got:
        assert [tuple(int(x) for x in r) for r in sampled] == [v.tup() for v in want_s], f"rank {rank}: sampled values"
        assert np.array_equal(quot, want_q), f"rank {rank}: quotient"
        if world > 1:
            assert timings["allgather_bytes_per_rank"] == 16 * (n_cols // world) * (world - 1)
