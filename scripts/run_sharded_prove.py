"""Sharded prove() on N GPUs (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/run_sharded_prove.py [log]
Every rank proves the same graphs through lb_prove_sharded; rank 0 checks the bytes against the single-GPU lb_prove and against
the committed CPU-prover fixtures, and prints timings + NCCL traffic as one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from luminair_b200.backend import CudaBackend
from luminair_b200.prover import Comm, last_stage_ms, prove, STAGE_NAMES
from luminair_b200.trace import DeviceGraphTrace
from luminair_b200.workloads import build_add_graph, build_wide, synthetic_add_graph_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
log = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
torch.cuda.set_device(local)
dist.init_process_group("gloo")  # only carries the 128-byte NCCL id and the barriers of this script
be = CudaBackend(local)
ids = [Comm.unique_id(be) if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = Comm(be, ids[0], rank, world)
out = {"world": world, "log": log}


def bench(name, rec, fixture):
    try:
        return _bench(name, rec, fixture)
    except Exception as e:  # show it: torchrun hides child tracebacks
        print(f"rank {rank}: {name}: {type(e).__name__}: {e}", file=sys.stderr, flush=True)
        raise


def _bench(name, rec, fixture):
    meta, dev, _ = rec.finish()
    single = prove(meta, backend=be, device_tables=dev)
    dist.barrier()
    sharded = prove(meta, backend=be, device_tables=dev, comm=comm)
    ts, t1 = [], []
    for _ in range(5):
        dist.barrier()
        t0 = time.perf_counter(); prove(meta, backend=be, device_tables=dev, comm=comm); ts.append((time.perf_counter() - t0) * 1e3)
    stages = last_stage_ms(be)
    stats = comm.stats()
    for _ in range(3):
        t0 = time.perf_counter(); prove(meta, backend=be, device_tables=dev); t1.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([min(ts)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fx = os.path.join(ROOT, "tests", "golden", fixture)
    eq_fix = (open(fx, "rb").read() == sharded) if (log == 20 and os.path.exists(fx)) else None
    out[name] = {"proof_equals_single_device": sharded == single, "proof_equals_cpu_prover_fixture": eq_fix,
                 "ms_sharded_max_over_ranks": float(t[0]), "ms_single_gpu": min(t1), "speedup": min(t1) / float(t[0]),
                 "stages_ms_rank": dict(zip(STAGE_NAMES, [round(x, 3) for x in stages])), "nccl": stats, "proof_bytes": len(sharded)}
    ok = torch.tensor([1 if sharded == single and eq_fix is not False else 0])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok[0])


a, b = synthetic_add_graph_inputs(log, seed=42)
ok1 = bench("cfg3_add", build_add_graph(DeviceGraphTrace(be), a, b), "cfg3_add_log20.proof.bin")
ok2 = bench("wide", build_wide(DeviceGraphTrace(be), log), "wide_log20.proof.bin")
if "--all-components" in sys.argv:
    # every component incl. the four lookup tables (test infrastructure: host tables and LUT layouts from the numpy checker),
    # non-default PcsConfig and the "v2" channel: sharded bytes == single-GPU bytes
    from luminair_b200.prover import PcsConfig
    from oracle import pie as opie
    ok3, shapes = True, {}
    # 2^16 elements: lookup tables (2^8 .. 2^15) no larger than their consumers; 2^11 elements: tables LARGER than the consumers'
    # traces and a blow-up of 4 - the constraint-framework `need_to_extend` path (columns re-sharded on the evaluation domain)
    for n_el, cases in ((1 << 16, ((None, "legacy"), (PcsConfig(7, 1, 1, 9), "v2"))),
                        (1 << 11, ((None, "legacy"), (PcsConfig(6, 2, 1, 5), "v2")))):
        pie, pre = opie.all_components_graph(n=n_el, seed=3)
        shapes[str(n_el)] = {k: list(v.shape) for k, v in pie}
        for cfg, variant in cases:
            single = prove(pie, backend=be, preprocessed=pre, config=cfg, channel_variant=variant)
            dist.barrier()
            sharded = prove(pie, backend=be, preprocessed=pre, config=cfg, channel_variant=variant, comm=comm)
            ok3 = ok3 and sharded == single
    flag = torch.tensor([1 if ok3 else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["all_components"] = {"proof_equals_single_device": bool(flag[0]), "tables": shapes,
                             "configs": ["2^16 elements: default / legacy channel; pow 7, last-layer bound 2, 9 queries / v2 channel",
                                         "2^11 elements (lookup tables larger than the traces): default; blow-up 4, pow 6, 5 queries / v2"]}
    ok2 = ok2 and bool(flag[0])
if rank == 0:
    print(json.dumps(out), flush=True)
comm.close()
be.close()
dist.destroy_process_group()
sys.exit(0 if (ok1 and ok2) else 1)
