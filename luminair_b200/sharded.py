"""Column-sharded commitment across GPUs (BASELINE cfg 5).

Trace columns are independent through interpolate / low-degree extension, so each rank owns a
contiguous range of columns and transforms them with no communication.  A Merkle leaf, however,
hashes *every* column of a row in column order (stwo ``MerkleOps::commit_on_layer`` as reached
from ``tree_builder.commit``, crates/prover/src/prover.rs:59,179,298), so a root that is
bit-identical to the single-device tree needs one real exchange:

  1. per rank:  interpolate + LDE of its columns                         (no collective)
  2. all-to-all: column shards -> row shards; rank r receives rows [r*R/W, (r+1)*R/W) of ALL columns
  3. per rank:  leaf layer + sub-tree over its row range                  (no collective)
  4. all-gather of the W sub-tree roots (32 B each), then every rank hashes the log2(W) top levels

One process per GPU, ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests) is
the plumbing; the kernels are the library's own (``ShardOps`` wraps the C ABI).  The host logic
below is backend-agnostic so that tests can drive it with a CPU stand-in for the kernels.
"""
from __future__ import annotations

import time
from typing import List, Optional, Protocol

import torch
import torch.distributed as dist


class ShardOps(Protocol):
    """The three device operations the sharded pipeline needs."""

    def lde(self, trace: torch.Tensor, log_size: int, log_blowup: int) -> torch.Tensor:
        """trace [n_cols, 2^log] (values, bit-reversed circle-domain order) -> evaluations [n_cols, 2^(log+blowup)]."""

    def merkle_layer(self, log_size: int, prev: Optional[torch.Tensor], cols: Optional[torch.Tensor]) -> torch.Tensor:
        """-> [2^log, 8] digests; prev [2^(log+1), 8] or None; cols [n_cols, 2^log] or None."""

    def eval_at_point(self, coeffs: torch.Tensor, log_size: int, point) -> torch.Tensor:
        """coeffs [n_cols, 2^log] -> sampled values [n_cols, 4] (QM31 coordinates); point = 8 u32 (x then y)."""

    def quotients_partial(self, lde: torch.Tensor, lde_log: int, point, values: torch.Tensor, random_coeff, col_offset: int,
                          n_cols_global: int) -> torch.Tensor:
        """Partial DEEP quotient [4, 2^lde_log] of this rank's columns (global column indices col_offset ..)."""

    def sync(self) -> None:
        ...


class CudaShardOps:
    """ShardOps over libluminair_b200 (device pointers of torch CUDA tensors cross the C ABI)."""

    def __init__(self, backend):
        self.be = backend

    def lde(self, trace, log_size, log_blowup):
        import ctypes as C
        from ._lib import check
        be = self.be
        n_cols = trace.shape[0]
        n = 1 << log_size
        assert trace.is_cuda and trace.dtype == torch.int32 and trace.is_contiguous() and trace.shape[1] == n
        torch.cuda.current_stream().synchronize()
        check(be.ctx, be.lib.lb_interpolate_batch(be.ctx, C.c_void_p(trace.data_ptr()), n, n_cols, log_size), "lb_interpolate_batch")
        out = torch.empty((n_cols, n << log_blowup), dtype=torch.int32, device=trace.device)
        check(be.ctx, be.lib.lb_evaluate_batch(be.ctx, C.c_void_p(trace.data_ptr()), n, log_size, C.c_void_p(out.data_ptr()),
                                                n << log_blowup, log_size + log_blowup, n_cols), "lb_evaluate_batch")
        return out

    def merkle_layer(self, log_size, prev, cols):
        n = 1 << log_size
        dev = prev.device if prev is not None else cols.device
        out = torch.empty((n, 8), dtype=torch.int32, device=dev)
        ptrs = [] if cols is None else [cols.data_ptr() + 4 * n * c for c in range(cols.shape[0])]
        if cols is not None:
            assert cols.is_contiguous() and cols.shape[1] == n
        self.be.merkle_commit_layer(log_size, prev.data_ptr() if prev is not None else None, ptrs, out.data_ptr())
        return out

    def eval_at_point(self, coeffs, log_size, point):
        import numpy as np
        n = 1 << log_size
        ptrs = [coeffs.data_ptr() + 4 * n * c for c in range(coeffs.shape[0])]
        out = self.be.eval_at_point(ptrs, log_size, list(point))
        return torch.from_numpy(out.view(np.int32).copy())

    def quotients_partial(self, lde, lde_log, point, values, random_coeff, col_offset, n_cols_global):
        import numpy as np
        n = 1 << lde_log
        n_cols = lde.shape[0]
        ptrs = [lde.data_ptr() + 4 * n * c for c in range(n_cols)]
        vals = values.numpy().view(np.uint32)
        out = torch.empty((4, n), dtype=torch.int32, device=lde.device)
        batch = (list(point), [(c, [int(x) for x in vals[c]]) for c in range(n_cols)])
        self.be.accumulate_quotients(lde_log, ptrs, [batch], random_coeff, [out.data_ptr() + 4 * n * k for k in range(4)],
                                     shards=[(col_offset, n_cols_global)])
        return out

    def sync(self):
        self.be.sync()


def column_range(n_cols_total: int, rank: int, world: int):
    """Contiguous column shard of `rank` (cfg 5: 256 columns / 8 ranks = 32 each)."""
    if n_cols_total % world:
        raise ValueError("the column count must be a multiple of the world size")
    per = n_cols_total // world
    return rank * per, (rank + 1) * per


def sharded_commit(ops: ShardOps, trace_local: torch.Tensor, log_size: int, log_blowup: int = 1, group=None,
                   timings: Optional[dict] = None) -> bytes:
    """Commit `world * n_cols_local` columns of 2^log_size rows, `n_cols_local` of them on each rank
    (rank r holds columns [r * n_cols_local, (r+1) * n_cols_local)).  Returns the 32-byte Merkle root,
    identical on every rank and identical to a single-device commit of all columns."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world & (world - 1):
        raise ValueError("world size must be a power of two")
    log_w = world.bit_length() - 1
    lde_log = log_size + log_blowup
    if lde_log < log_w:
        raise ValueError("fewer rows than ranks")
    n_cols_local = trace_local.shape[0]
    rows_local = (1 << lde_log) // world
    t0 = time.perf_counter()

    # 1. transform own columns
    lde = ops.lde(trace_local, log_size, log_blowup)  # [n_cols_local, 2^lde_log]
    ops.sync()
    t1 = time.perf_counter()

    # 2. column shards -> row shards
    if world > 1:
        send = lde.view(n_cols_local, world, rows_local).transpose(0, 1).contiguous()  # [world, n_cols_local, rows_local]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        cols = recv.view(world * n_cols_local, rows_local)  # all columns, global order, my rows
        del send
    else:
        cols = lde
    if cols.is_cuda:
        torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()

    # 3. sub-tree over my rows
    sub_log = lde_log - log_w
    layer = ops.merkle_layer(sub_log, None, cols)
    for lg in range(sub_log - 1, -1, -1):
        layer = ops.merkle_layer(lg, layer, None)
    ops.sync()
    t3 = time.perf_counter()

    # 4. gather the sub-tree roots, hash the top levels everywhere
    if world > 1:
        roots = torch.empty((world, 8), dtype=torch.int32, device=layer.device)
        dist.all_gather_into_tensor(roots, layer.reshape(1, 8).contiguous(), group=group)
        if roots.is_cuda:
            torch.cuda.current_stream().synchronize()
        layer = roots
        for lg in range(log_w - 1, -1, -1):
            layer = ops.merkle_layer(lg, layer, None)
        ops.sync()
    root = layer.reshape(-1).cpu().numpy().astype("<u4").tobytes()
    t4 = time.perf_counter()
    if timings is not None:
        timings.update({"lde_ms": (t1 - t0) * 1e3, "all_to_all_ms": (t2 - t1) * 1e3, "subtree_ms": (t3 - t2) * 1e3,
                        "root_allgather_ms": (t4 - t3) * 1e3, "total_ms": (t4 - t0) * 1e3,
                        "all_to_all_bytes_per_rank": int(n_cols_local * (1 << lde_log) * 4 * (world - 1) // world),
                        "rank": rank, "world": world})
    return root


M31_P = (1 << 31) - 1


def sharded_quotient_accumulation(ops: ShardOps, coeffs_local: torch.Tensor, lde_local: torch.Tensor, log_size: int,
                                  log_blowup: int, point, random_coeff, group=None, timings: Optional[dict] = None):
    """OODS sampling + DEEP quotient accumulation of a column-sharded commitment (SURVEY 8e collectives (2), (3)):

      1. per rank:   eval_at_point of its own polynomials                                     (no collective)
      2. all-gather: the sampled values, 16 B per column (they are mixed into the channel in global column order)
      3. per rank:   partial quotient of its own columns, weighted with the GLOBAL powers of the random coefficient
      4. all-reduce: coordinate-wise sum of the partial quotients (int64 sum, then mod p)

    `point` = 8 u32 (QM31 x, y), `random_coeff` = 4 u32.  Returns (sampled [n_cols_total, 4], quotient [4, 2^lde_log]),
    both identical on every rank and identical to the single-device ``eval_at_point`` / ``accumulate_quotients``
    over all columns (every column sampled at `point` only, as in a synthetic trace without LogUp columns)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_local = coeffs_local.shape[0]
    lde_log = log_size + log_blowup
    t0 = time.perf_counter()
    sampled_local = ops.eval_at_point(coeffs_local, log_size, point)  # [n_local, 4] on the host
    ops.sync()
    t1 = time.perf_counter()
    if world > 1:
        dev = lde_local.device
        mine = sampled_local.to(dev)
        gathered = torch.empty((world * n_local, 4), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(gathered, mine, group=group)
        sampled = gathered.cpu()
    else:
        sampled = sampled_local
    t2 = time.perf_counter()
    partial = ops.quotients_partial(lde_local, lde_log, point, sampled_local, random_coeff, rank * n_local, world * n_local)
    ops.sync()
    t3 = time.perf_counter()
    if world > 1:
        acc = partial.to(torch.int64)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        quotient = (acc % M31_P).to(torch.int32)
        if quotient.is_cuda:
            torch.cuda.current_stream().synchronize()
    else:
        quotient = partial
    t4 = time.perf_counter()
    if timings is not None:
        timings.update({"sample_ms": (t1 - t0) * 1e3, "sample_allgather_ms": (t2 - t1) * 1e3, "quotients_ms": (t3 - t2) * 1e3,
                        "quotient_allreduce_ms": (t4 - t3) * 1e3, "total_ms": (t4 - t0) * 1e3,
                        "allgather_bytes_per_rank": int(16 * n_local * (world - 1)),
                        "allreduce_bytes_per_rank": int(8 * 4 * (1 << lde_log)) if world > 1 else 0})
    return sampled, quotient


class FusedShardedCommitter:
    """The same commitment with the all-to-all FUSED into the transform: the last pass of each rank's LDE stores every
    4096-row tile straight into the receive buffer of the rank that owns those rows (CUDA IPC peer memory over
    NVLink / NVSwitch), already in [column][local row] layout.  No pack copy, no NCCL data-path call; the transfer
    overlaps the butterflies tile by tile.  torch.distributed only carries the 64-byte IPC handles (once), a barrier
    and the 32-byte sub-tree roots.

    setup() allocates and exchanges the buffers once; commit() can then be called repeatedly."""

    def __init__(self, backend, n_cols_local: int, log_size: int, log_blowup: int = 1, group=None):
        self.be, self.group = backend, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world & (self.world - 1) or self.world > 8:
            raise ValueError("world size must be a power of two <= 8")
        self.log_w = self.world.bit_length() - 1
        self.log_size, self.lde_log = log_size, log_size + log_blowup
        if self.lde_log < 16 or self.lde_log - self.log_w < 12:
            raise ValueError("fused path needs 2^16 LDE rows and 4096 rows per rank")
        self.n_cols_local = n_cols_local
        self.rows_local = (1 << self.lde_log) >> self.log_w
        self.recv = None
        self.peers: List[int] = []
        self._opened: List[int] = []

    def setup(self):
        import ctypes as C
        from ._lib import check
        be = self.be
        total_cols = self.n_cols_local * self.world
        self.recv = be.alloc(total_cols * self.rows_local)          # all columns x my rows
        self.scratch = be.alloc(self.n_cols_local << self.lde_log)  # intermediate passes of my columns
        handle = C.create_string_buffer(64)
        check(be.ctx, be.lib.lb_ipc_export(be.ctx, C.c_void_p(self.recv.ptr), handle), "lb_ipc_export")
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, handle.raw, group=self.group)
        else:
            handles = [handle.raw]
        self.peers = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.peers.append(self.recv.ptr)
                continue
            p = C.c_void_p()
            check(be.ctx, be.lib.lb_ipc_open(be.ctx, h, C.byref(p)), "lb_ipc_open")
            self.peers.append(p.value)
            self._opened.append(p.value)
        # tree layers over my rows + the top levels
        sub_log = self.lde_log - self.log_w
        self.layers = [be.alloc(8 << lg) for lg in range(sub_log + 1)]
        self.top = [be.alloc(8 << lg) for lg in range(self.log_w + 1)]

    def close(self):
        import ctypes as C
        for p in self._opened:
            self.be.lib.lb_ipc_close(self.be.ctx, C.c_void_p(p))
        self._opened = []

    def commit(self, trace_ptr: int, timings: Optional[dict] = None) -> bytes:
        """trace_ptr: device pointer of my n_cols_local x 2^log_size trace values (overwritten by the coefficients)."""
        import ctypes as C
        import numpy as np
        from ._lib import check
        be, W = self.be, self.world
        n = 1 << self.log_size
        t0 = time.perf_counter()
        check(be.ctx, be.lib.lb_interpolate_batch(be.ctx, C.c_void_p(trace_ptr), n, self.n_cols_local, self.log_size), "lb_interpolate_batch")
        peers = (C.c_void_p * W)(*self.peers)
        check(be.ctx, be.lib.lb_evaluate_batch_scatter(be.ctx, C.c_void_p(trace_ptr), n, self.log_size, C.c_void_p(self.scratch.ptr),
                                                        1 << self.lde_log, self.lde_log, self.n_cols_local, peers, W,
                                                        self.rank * self.n_cols_local), "lb_evaluate_batch_scatter")
        be.sync()
        if W > 1:
            dist.barrier(group=self.group)  # every rank's remote stores have landed
        t1 = time.perf_counter()
        sub_log = self.lde_log - self.log_w
        cols = [self.recv.ptr + 4 * self.rows_local * c for c in range(self.n_cols_local * W)]
        be.merkle_commit_layer(sub_log, None, cols, self.layers[sub_log].ptr)
        for lg in range(sub_log - 1, -1, -1):
            be.merkle_commit_layer(lg, self.layers[lg + 1].ptr, [], self.layers[lg].ptr)
        my_root = be.download(self.layers[0])
        t2 = time.perf_counter()
        if W > 1:
            mine = torch.from_numpy(my_root.view(np.int32).copy()).cuda().reshape(1, 8)
            roots = torch.empty((W, 8), dtype=torch.int32, device=mine.device)
            dist.all_gather_into_tensor(roots, mine, group=self.group)
            be.upload(roots.cpu().numpy().view(np.uint32).reshape(-1), self.top[self.log_w])
            for lg in range(self.log_w - 1, -1, -1):
                be.merkle_commit_layer(lg, self.top[lg + 1].ptr, [], self.top[lg].ptr)
            root = be.download(self.top[0]).astype("<u4").tobytes()
        else:
            root = my_root.astype("<u4").tobytes()
        t3 = time.perf_counter()
        if timings is not None:
            timings.update({"lde_scatter_ms": (t1 - t0) * 1e3, "subtree_ms": (t2 - t1) * 1e3, "root_allgather_ms": (t3 - t2) * 1e3,
                            "total_ms": (t3 - t0) * 1e3,
                            "nvlink_store_bytes_per_rank": int(self.n_cols_local * (1 << self.lde_log) * 4 * (W - 1) // W)})
        return root
