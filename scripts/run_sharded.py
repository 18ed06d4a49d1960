"""torchrun entry: column-sharded commit across N GPUs, checked against a single-GPU commit of all columns.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/run_sharded.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from luminair_b200.backend import CudaBackend
from luminair_b200.sharded import CudaShardOps, column_range, sharded_commit, sharded_quotient_accumulation

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
be = CudaBackend(lr)
log = int(os.environ.get("LOG", 16)); n_cols = int(os.environ.get("NCOLS", 64))
rng = np.random.Generator(np.random.PCG64(5))
full = rng.integers(0, (1 << 31) - 1, size=(n_cols, 1 << log), dtype=np.uint64).astype(np.uint32)
lo, hi = column_range(n_cols, rank, world)
tm = {}
root = sharded_commit(CudaShardOps(be), torch.from_numpy(full[lo:hi].view(np.int32).copy()).cuda(), log, 1, timings=tm)
# reference: everything on this rank's GPU as one shard (world-size-1 code path, no collectives)
class _Solo:  # run the same function outside the process group
    pass
lde = CudaShardOps(be).lde(torch.from_numpy(full.view(np.int32).copy()).cuda(), log, 1)
ops = CudaShardOps(be)
layer = ops.merkle_layer(log + 1, None, lde)
for lg in range(log, -1, -1):
    layer = ops.merkle_layer(lg, layer, None)
be.sync()
want = layer.reshape(-1).cpu().numpy().astype("<u4").tobytes()
# the fused variant: all-to-all done by the last CFFT pass through peer memory
from luminair_b200.sharded import FusedShardedCommitter
fc = FusedShardedCommitter(be, hi - lo, log, 1)
fc.setup()
ftm = {}
for it in range(3):
    tr = be.upload(full[lo:hi].reshape(-1))
    be.sync(); dist.barrier()
    froot = fc.commit(tr.ptr, timings=ftm)
print(f"rank {rank}/{world}: FUSED root {froot.hex()[:16]} equal={froot == want} total {ftm['total_ms']:.2f} ms "
      f"(lde+scatter {ftm['lde_scatter_ms']:.2f}, subtree {ftm['subtree_ms']:.2f}, roots {ftm['root_allgather_ms']:.2f})", flush=True)
assert froot == want
dist.barrier()
fc.close()
dist.barrier()
print(f"rank {rank}/{world}: sharded root {root.hex()[:16]} single-device root {want.hex()[:16]} equal={root == want} "
      f"total {tm['total_ms']:.2f} ms (lde {tm['lde_ms']:.2f}, a2a {tm['all_to_all_ms']:.2f}, subtree {tm['subtree_ms']:.2f})", flush=True)
assert root == want
# OODS sampling + DEEP quotient accumulation over the column shards vs one device holding every column
POINT, RC = [11, 22, 33, 44, 55, 66, 77, 88], [5, 6, 7, 8]
n = 1 << log
coeffs_local = torch.from_numpy(full[lo:hi].view(np.int32).copy()).cuda()
lde_local = ops.lde(coeffs_local, log, 1)  # coeffs_local now holds the coefficients
qtm = {}
sampled, quot = sharded_quotient_accumulation(ops, coeffs_local, lde_local, log, 1, POINT, RC, timings=qtm)
coeffs_all = torch.from_numpy(full.view(np.int32).copy()).cuda()
lde_all = ops.lde(coeffs_all, log, 1)
want_s = ops.eval_at_point(coeffs_all, log, POINT)
want_q = ops.quotients_partial(lde_all, log + 1, POINT, want_s, RC, 0, n_cols)
be.sync()
ok_s = bool(torch.equal(sampled, want_s)); ok_q = bool(torch.equal(quot.cpu(), want_q.cpu()))
print(f"rank {rank}/{world}: sharded sampled values equal={ok_s} quotient equal={ok_q} total {qtm['total_ms']:.2f} ms "
      f"(sample {qtm['sample_ms']:.2f}, allgather {qtm['sample_allgather_ms']:.2f}, quotients {qtm['quotients_ms']:.2f}, "
      f"allreduce {qtm['quotient_allreduce_ms']:.2f})", flush=True)
assert ok_s and ok_q
dist.destroy_process_group()
