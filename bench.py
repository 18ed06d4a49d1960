#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json configs[1]):

  single CFFT: 2^20-row x 64 M31 columns, inverse (interpolate) + forward (evaluate)
  round trip per step, on 1 GPU; N GPUs = N independent column shards (weak scaling,
  no data-path collective — columns are independent).

Prints ONE JSON line (see the contract in the task description).  `--impl reference`
times the CPU restatement of the same path (oracle/c, OpenMP over columns) instead:
the reference itself (Rust + un-vendored stwo) cannot be built in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCU_TRAFFIC_BYTES_PER_LAUNCH = 486.0e6  # measured once with ncu --set full (profiles/), not re-measured by this script
LOG_N = 20
N_COLS = 64
SEED = 20260101
P = (1 << 31) - 1


def field_ops_per_step(log_n=LOG_N, n_cols=N_COLS):
    n = 1 << log_n
    # (N/2) log N butterflies x 3 ops, two transforms, + N scaling mults on interpolate
    return 2 * n_cols * 3 * (n // 2) * log_n + n_cols * n


def algorithmic_bytes_per_step(log_n=LOG_N, n_cols=N_COLS):
    return 2 * 8 * (1 << log_n) * n_cols  # 8 B / element / column / transform


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def gen_inputs(n_cols, log_n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, P, size=(n_cols, 1 << log_n), dtype=np.uint64).astype(np.uint32)


# ---------------------------------------------------------------------------------------
def cpu_roundtrip(sample_cols, threads, reps=3):
    """Time the packed (AVX-512 / AVX2) + OpenMP CFFT of the CPU prover (oracle/c/cpu_prover) on `sample_cols` columns.
    Returns seconds per interpolate + evaluate round trip (best of `reps`)."""
    from oracle import cpu_prover as cp
    v = gen_inputs(sample_cols, LOG_N, SEED)
    ref = v.copy()
    cp.cfft(v, forward=False, n_threads=threads)  # warm (page-in, thread pool, twiddles)
    cp.cfft(v, forward=True, n_threads=threads)
    if not np.array_equal(v, ref):
        raise SystemExit("bench: CPU CFFT round trip did not reproduce its input")
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        cp.cfft(v, forward=False, n_threads=threads)
        cp.cfft(v, forward=True, n_threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def cpu_prove_baseline(log, threads, gpu_proof=None, wide=True):
    """The compiled CPU prover (oracle/c/cpu_prover: packed M31 arithmetic, OpenMP; the restatement of the reference's
    SimdBackend + rayon path) on the WHOLE cfg-3 workload (and the 2^log x 61-column wide trace), all host cores."""
    from oracle import cpu_prover as cp
    from oracle import pie as opie
    pie = opie.synthetic_add_graph_pie(log, seed=42)
    cp.prove(pie, n_threads=threads)  # warm: domain tables, buffer pool
    ts, stages, proof = [], None, None
    for _ in range(3):
        t0 = time.perf_counter()
        proof, st = cp.prove(pie, n_threads=threads, return_stages=True)
        dt = (time.perf_counter() - t0) * 1e3
        if not ts or dt < min(ts):
            stages = st
        ts.append(dt)
    out = {"value": min(ts), "unit": "ms per proof", "median_ms": statistics.median(ts), "proofs_per_s": 1e3 / min(ts),
           "cores": threads, "kind": "port", "lanes": cp.lanes(),
           "sample": f"the whole GPU workload: a+b graph at 2^{log} elements (Add 2^{log} x 15 + Inputs 2^{log + 1} x 7), "
                     f"{threads} threads, {cp.lanes()}-lane packed M31 (oracle/c/cpu_prover, C++/OpenMP); a restatement of the "
                     "reference's SimdBackend path, not stwo itself (Rust toolchain absent)",
           "stages_ms": {k: round(v, 2) for k, v in stages.items()}}
    if gpu_proof is not None:
        out["proof_bytes_equal_gpu"] = proof == gpu_proof
    if wide:
        wpie = opie.wide_graph(log)
        cp.prove(wpie, n_threads=threads)
        t0 = time.perf_counter()
        cp.prove(wpie, n_threads=threads)
        out["wide_ms_per_proof"] = (time.perf_counter() - t0) * 1e3
        out["wide_sample"] = f"the whole wide workload (2^{log} rows x 61 main-trace columns + Inputs), one timed proof after one warm-up"
    return out


def run_reference(args):
    """The reference arm: the CPU implementation of the same path (the packed + OpenMP CFFT of oracle/c/cpu_prover; the Rust
    reference cannot be built here) on ALL host cores - the thread count is taken from the affinity mask, not from
    OMP_NUM_THREADS (torchrun sets that to 1) - on the same 64 x 2^20 round trip per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_prover as cp
    threads = cp.host_cores()
    sample_cols = N_COLS  # the whole per-GPU workload every step (about 0.1-0.2 s of CPU work per step)
    v = gen_inputs(sample_cols, LOG_N, SEED)
    for _ in range(max(args.warmup, 1)):
        cp.cfft(v, forward=False, n_threads=threads)
        cp.cfft(v, forward=True, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cp.cfft(v, forward=False, n_threads=threads)
        cp.cfft(v, forward=True, n_threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    ops = field_ops_per_step(LOG_N, sample_cols)
    value = ops / dt
    line = {
        "impl": "reference",
        "metric": "cfft_roundtrip_m31_field_ops_per_s", "value": value, "unit": "M31 field-ops/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (M31)", "data": "synthetic",
        "config": {"workload": "cfft_roundtrip 2^20 rows x 64 M31 columns per GPU (BASELINE configs[1])",
                   "log_rows": LOG_N, "n_cols_per_gpu": N_COLS},
        "cpu_baseline": {"value": value, "unit": "M31 field-ops/s", "cores": threads, "kind": "port", "lanes": cp.lanes(),
                         "sample": f"all {N_COLS} columns x 2^{LOG_N} per step, interpolate+evaluate, {threads} threads "
                                   f"(affinity mask; OMP_NUM_THREADS ignored), {cp.lanes()}-lane packed M31 butterflies, "
                                   "cache-blocked; CPU restatement (oracle/c/cpu_prover), not stwo SimdBackend (Rust toolchain absent)"},
        "e2e": {"value": value, "unit": "M31 field-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def pin_to_gpu_numa_node(index):
    """Bind this rank to the CPUs NVML reports as local to its GPU, so the pinned host buffers of the e2e leg are
    first-touched on that NUMA node (8 ranks otherwise share one node's memory controllers).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to GPU {index}"
    except Exception as e:  # containers may forbid it
        return f"not applied ({type(e).__name__})"
    return "not applied"


# ---------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = pin_to_gpu_numa_node(local_rank)  # before any pinned allocation: first touch lands on the GPU's NUMA node
    # stdout carries the one JSON line only: whatever native libraries write to fd 1 on the way (the "NCCL version ..." banner
    # of the first collective, NCCL_DEBUG output) is sent to stderr; the saved descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from luminair_b200.backend import ColumnBatch, CudaBackend
    be = CudaBackend(local_rank)
    n = 1 << LOG_N
    # each rank owns its own 64-column shard (weak scaling; different seed per rank)
    host = torch.empty((N_COLS, n), dtype=torch.int32, pin_memory=True)
    host_np = host.numpy().view(np.uint32)
    host_np[:] = gen_inputs(N_COLS, LOG_N, SEED + rank)
    host_out = torch.empty((N_COLS, n), dtype=torch.int32, pin_memory=True)
    out_np = host_out.numpy().view(np.uint32)

    d_in = be.alloc(N_COLS * n)
    cols = ColumnBatch(d_in, N_COLS, LOG_N)
    be.precompute_twiddles(LOG_N)
    lib, ctx = be.lib, be.ctx

    def h2d():
        lib.lb_upload(ctx, C.c_void_p(d_in.ptr), C.c_void_p(host_np.ctypes.data), N_COLS * n)

    def step():
        # in place: the round trip restores the column values, so every step sees the same input
        be.interpolate(cols)          # values -> coefficients
        be.evaluate(cols, cols)       # coefficients -> values

    def d2h():
        lib.lb_download(ctx, C.c_void_p(out_np.ctypes.data), C.c_void_p(d_in.ptr), N_COLS * n)

    def barrier():
        be.sync()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()

    # ---- correctness guard (cheap): round trip must return the input
    h2d()
    step()
    d2h()
    if not np.array_equal(out_np, host_np):
        raise SystemExit("bench: CFFT round trip did not reproduce its input")

    # ---- device-resident timing (inputs already in HBM), CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    be.timer_start()
    for _ in range(args.steps):
        step()
    total_ms = be.timer_stop_ms()
    barrier()
    # keep the GPU under the same load while nvidia-smi samples (a 15 ms timed region is shorter than one sample period):
    # the clocks reported are those of a >= 1 s run of the very same step
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 1.2:
        for _ in range(50):
            step()
        be.sync()
    clocks = sampler.stop()

    # per-kernel (dominant: cfft_pass_kernel) timing: interpolate and evaluate separately
    t_int = t_ev = 0.0
    reps = 5
    for _ in range(reps):
        be.timer_start(); be.interpolate(cols); t_int += be.timer_stop_ms() / reps
        be.timer_start(); be.evaluate(cols, cols); t_ev += be.timer_stop_ms() / reps

    # ---- end-to-end through the C ABI with HOST buffers (pinned): lb_lde_host = upload + interpolate + evaluate +
    # download per step, chunked over three streams inside the library so copies and transforms overlap
    def e2e_step():
        be.lde_host(host_np, out=out_np)

    e2e_s = None
    if not args.no_e2e:  # --no-e2e: the ncu launch list of this command then holds the timed 64-column launches only
        e2e_step()
        if not np.array_equal(out_np, host_np):
            raise SystemExit("bench: lb_lde_host round trip did not reproduce its input")
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps

    # ---- full prove() on the BASELINE cfg-3 shape (Add 2^20 rows + Inputs 2^21 rows), rank-local
    prove_info = None
    if not args.no_prove:
        prove_info = bench_prove(be, torch, args)
        if not distributed:
            prove_info["concurrent"] = bench_concurrent_prove(args, prove_info["_proof_bytes"])

    # ---- N > 1: the column-sharded commit of BASELINE cfg 5 (the one place the path has a real exchange)
    sharded_info = None
    sharded_prove_info = None
    if distributed and not args.no_sharded:
        sharded_prove_info = bench_sharded_prove(be, torch, dist, args, rank, world)
        sharded_info = bench_sharded_commit(be, torch, dist, args, rank, world, local_rank)

    # ---- reduce over ranks (max time)
    t = torch.tensor([total_ms, e2e_s if e2e_s is not None else 0.0], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    ops = field_ops_per_step() * world
    value = ops / (ms_per_step * 1e-3)
    e2e_value = ops / e2e_s if e2e_s else None

    if rank == 0:
        peak, peak_kind = load_peaks()
        n_launch_per_step = 4  # 2 passes per transform at log 20
        alg_bytes_launch = algorithmic_bytes_per_step() / n_launch_per_step
        # the timed region is K steps of exactly these four launches, back to back on one stream, bracketed by CUDA events:
        # its per-launch average is the duration the roofline uses (the separately timed interpolate / evaluate calls below
        # carry one event pair per call and are reported for the split only)
        avg_launch_ms = ms_per_step / n_launch_per_step
        achieved = alg_bytes_launch / (avg_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "cfft_roundtrip_m31_field_ops_per_s", "value": value, "unit": "M31 field-ops/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (M31)",
            "data": "synthetic",
            "config": {"workload": "cfft_roundtrip 2^20 rows x 64 M31 columns per GPU (BASELINE configs[1])",
                       "log_rows": LOG_N, "n_cols_per_gpu": N_COLS, "l2": "working set 256 MiB per GPU, larger than L2 (126 MB); in place",
                       "sharding": "columns, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "M31 field-ops/s", "h2d_bytes_per_step": N_COLS * n * 4 * world,
                    "d2h_bytes_per_step": N_COLS * n * 4 * world, "ms_per_step": e2e_s * 1e3 if e2e_s else None,
                    "host_numa_binding": numa},
            "gpu_launches": args.steps * n_launch_per_step,
            "roofline": {"bound": "hbm", "kernel": "cfft_low_fast / cfft_high_vec (the 4 CFFT passes of a step)", "achieved": achieved, "peak": peak,
                         "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum, mean of the 4 passes, profiles/r2_cfft_ncu_full_summary.csv (same kernels as profiles/r1_cfft_v8_ncu_full_summary.csv)",
                         "algorithmic_bytes_per_launch": alg_bytes_launch, "avg_launch_ms": avg_launch_ms,
                         "interpolate_ms": t_int, "evaluate_ms": t_ev},
        }
        if prove_info:
            line["prove"] = prove_info
        if sharded_prove_info:
            line["sharded_prove"] = sharded_prove_info
        if sharded_info:
            line["sharded_commit"] = sharded_info
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again, as in --impl reference
            threads = len(all_cpus)
            sample_cols = N_COLS
            dt = cpu_roundtrip(sample_cols, threads)
            from oracle import cpu_prover as _cp
            line["cpu_baseline"] = {"value": field_ops_per_step(LOG_N, sample_cols) / dt, "unit": "M31 field-ops/s",
                                    "cores": threads, "kind": "port", "lanes": _cp.lanes(),
                                    "sample": f"all {N_COLS} columns x 2^{LOG_N}, interpolate+evaluate round trip, best of 3, "
                                              f"{_cp.lanes()}-lane packed M31 + OpenMP, cache-blocked (oracle/c/cpu_prover)"}
            if prove_info:
                cpu_p = cpu_prove_baseline(args.prove_log, threads, gpu_proof=prove_info.pop("_proof_bytes", None))
                line["cpu_baseline"]["prove"] = cpu_p
                prove_info["speedup_vs_cpu_prover"] = {
                    "device_resident": cpu_p["value"] / prove_info["ms_device_resident"]["median"],
                    "e2e_host_tensors": cpu_p["value"] / prove_info["ms_e2e_host_tensors"]["median"],
                    "cpu_cores": threads,
                    "note": "proofs/s of lb_prove on 1 x B200 over proofs/s of the compiled CPU prover on all host cores, same "
                            "2^%d workload, proof bytes equal: %s (north-star target >= 10x)" % (args.prove_log, cpu_p.get("proof_bytes_equal_gpu"))}
        if prove_info:
            prove_info.pop("_proof_bytes", None)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()
    be.close()


def bench_sharded_prove(be, torch, dist, args, rank, world):
    """prove() of ONE proof by all N GPUs together (lb_prove_sharded: strong scaling; the total work is that of the 1-GPU
    proof): BASELINE configs[2] (a + b at 2^log) and the 2^log x 61-column wide trace.  Every rank checks that the bytes equal
    its own single-GPU lb_prove of the same tables and, at log 20, the committed CPU-prover fixture."""
    from luminair_b200.prover import STAGE_NAMES, Comm, last_stage_ms, prove
    from luminair_b200.trace import DeviceGraphTrace
    from luminair_b200.workloads import build_add_graph, build_wide, synthetic_add_graph_inputs
    log = args.prove_log
    ids = [Comm.unique_id(be) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = Comm(be, ids[0], rank, world)
    out = {"how": "lb_prove_sharded: column-owned interpolate / LDE / sampling, one grouped NCCL exchange per tree into row shards, "
                  "row-sharded Merkle sub-trees (all-gather of 32-byte roots), constraint quotients, DEEP quotients and FRI layers; "
                  "replicated Fiat-Shamir channel; device-resident (replicated) trace tables",
           "n_gpus": world, "scaling": "strong"}
    a, b = synthetic_add_graph_inputs(log, seed=42)
    for name, rec, fixture in (("cfg3_add", build_add_graph(DeviceGraphTrace(be), a, b), "cfg3_add_log20.proof.bin"),
                               ("wide", build_wide(DeviceGraphTrace(be), log), "wide_log20.proof.bin")):
        meta, dev, _ = rec.finish()
        single = prove(meta, backend=be, device_tables=dev)
        t1 = []
        for _ in range(3):
            t0 = time.perf_counter()
            prove(meta, backend=be, device_tables=dev)
            t1.append((time.perf_counter() - t0) * 1e3)
        dist.barrier()
        sharded = prove(meta, backend=be, device_tables=dev, comm=comm)
        ts = []
        for _ in range(6):
            be.sync()
            dist.barrier()
            t0 = time.perf_counter()
            prove(meta, backend=be, device_tables=dev, comm=comm)
            ts.append((time.perf_counter() - t0) * 1e3)
        stages, stats = last_stage_ms(be), comm.stats()
        t = torch.tensor([min(ts), statistics.median(ts), -min(t1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # slowest rank of the sharded proof, fastest single-GPU proof
        fx = os.path.join(ROOT, "tests", "golden", fixture)
        eq_fix = (open(fx, "rb").read() == sharded) if (log == 20 and os.path.exists(fx)) else None
        ok = torch.tensor([1 if (sharded == single and eq_fix is not False) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        ms_n, ms_1 = float(t[0]), -float(t[2])
        out[name] = {"proof_equals_single_device": bool(ok[0]), "proof_equals_cpu_prover_fixture": eq_fix,
                     "ms_per_proof_max_over_ranks": {"min": ms_n, "median": float(t[1])}, "ms_per_proof_1gpu": ms_1,
                     "speedup_vs_1gpu": ms_1 / ms_n, "strong_scaling_efficiency": ms_1 / ms_n / world,
                     "stages_ms_rank0": dict(zip(STAGE_NAMES, [round(x, 3) for x in stages])),
                     "nccl_bytes_sent_per_rank": stats["bytes_sent"], "nccl_bytes_received_per_rank": stats["bytes_received"],
                     "nccl_grouped_calls": stats["n_collectives"], "proof_bytes": len(sharded),
                     "timer": "host wall clock around lb_prove_sharded after a barrier, max over ranks"}
        if not bool(ok[0]):
            raise SystemExit(f"bench: sharded proof of {name} differs from the single-GPU proof")
        del rec
    comm.close()
    return out


def bench_sharded_commit(be, torch, dist, args, rank, world, local_rank):
    """BASELINE cfg 5 per-GPU share: 32 columns x 2^22 rows per rank (256 columns at 8 GPUs), interpolate + LDE
    (no collective) -> NCCL all-to-all to row shards -> sub-tree hashing -> all-gather of the sub-tree roots."""
    from luminair_b200.sharded import CudaShardOps, sharded_commit
    log, ncols = args.sharded_log, args.sharded_cols_per_gpu
    ops = CudaShardOps(be)
    dev = torch.device("cuda", local_rank)
    g = torch.Generator(device=dev)
    g.manual_seed(5 + rank)
    base = torch.randint(0, P, (ncols, 1 << log), dtype=torch.int32, device=dev, generator=g)
    best, root = None, None
    for it in range(3):
        trace = base.clone()
        torch.cuda.synchronize()
        dist.barrier()
        tm = {}
        root = sharded_commit(ops, trace, log, 1, timings=tm)
        t = torch.tensor([tm["total_ms"], tm["lde_ms"], tm["all_to_all_ms"], tm["subtree_ms"], tm["root_allgather_ms"]],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0 and (best is None or float(t[0]) < best[0]):
            best = [float(x) for x in t]
            a2a = tm["all_to_all_bytes_per_rank"]
        del trace
    # fused variant: the all-to-all is done by the last CFFT pass storing into NVLink peer memory (no NCCL data-path call)
    fused = None
    try:
        from luminair_b200.sharded import FusedShardedCommitter
        fc = FusedShardedCommitter(be, ncols, log, 1)
        fc.setup()
        fbest = None
        froot = None
        for it in range(4):
            tr = be.alloc(ncols << log)
            be.lib.lb_copy(be.ctx, C.c_void_p(tr.ptr), C.c_void_p(base.data_ptr()), ncols << log)
            be.sync()
            torch.cuda.synchronize()
            dist.barrier()
            tm = {}
            froot = fc.commit(tr.ptr, timings=tm)
            t = torch.tensor([tm["total_ms"], tm["lde_scatter_ms"], tm["subtree_ms"], tm["root_allgather_ms"]],
                             dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it > 0 and (fbest is None or float(t[0]) < fbest[0]):
                fbest = [float(x) for x in t]
            tr.free()
        dist.barrier()
        fc.close()
        dist.barrier()  # nobody frees an exported buffer while a peer still has it mapped
        fused = {"ms_total_max_over_ranks": fbest[0], "ms_lde_with_peer_scatter": fbest[1], "ms_subtree": fbest[2],
                 "ms_root_allgather_and_top": fbest[3], "nvlink_store_bytes_per_rank": tm["nvlink_store_bytes_per_rank"],
                 "root_equals_nccl_variant": froot == root,
                 "how": "last CFFT pass stores each 4096-row tile into the owner rank's buffer (CUDA IPC peer memory); no pack copy, no NCCL all-to-all"}
    except Exception as e:  # report, do not hide
        fused = {"error": repr(e)}
    # OODS sampling + DEEP quotient accumulation over the column shards (SURVEY 8e collectives (2), (3))
    quot = None
    try:
        from luminair_b200.sharded import sharded_quotient_accumulation
        coeffs = base.clone()
        lde = ops.lde(coeffs, log, 1)
        qbest = None
        for it in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            tm = {}
            sharded_quotient_accumulation(ops, coeffs, lde, log, 1, [11, 22, 33, 44, 55, 66, 77, 88], [5, 6, 7, 8], timings=tm)
            t = torch.tensor([tm["total_ms"], tm["sample_ms"], tm["sample_allgather_ms"], tm["quotients_ms"],
                              tm["quotient_allreduce_ms"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it > 0 and (qbest is None or float(t[0]) < qbest[0]):
                qbest = [float(x) for x in t]
        quot = {"ms_total_max_over_ranks": qbest[0], "ms_eval_at_point": qbest[1], "ms_sample_allgather": qbest[2],
                "ms_partial_quotients": qbest[3], "ms_quotient_allreduce": qbest[4],
                "nccl_allgather_bytes_per_rank": tm["allgather_bytes_per_rank"],
                "nccl_allreduce_bytes_per_rank": tm["allreduce_bytes_per_rank"],
                "how": "each rank samples and accumulates its own columns with the global random-coefficient powers "
                       "(lb_accumulate_quotients_shard); partial quotients are summed with one int64 all-reduce"}
        del coeffs, lde
    except Exception as e:  # report, do not hide
        quot = {"error": repr(e)}
    return {"fused_all_to_all": fused, "quotient_accumulation": quot,
            "workload": f"column-sharded commit: {ncols} columns x 2^{log} rows per GPU ({ncols * world} columns total), blow-up 2, "
                        "interpolate + LDE -> NCCL all-to-all (columns -> rows) -> Blake2s sub-trees -> all-gather of roots "
                        "(BASELINE configs[4] shape; root bit-identical to a single-device tree, tests/test_sharded_gloo.py)",
            "ms_total_max_over_ranks": best[0], "ms_lde": best[1], "ms_all_to_all": best[2], "ms_subtree": best[3],
            "ms_root_allgather_and_top": best[4], "nccl_all_to_all_bytes_per_rank": a2a, "root": root.hex(),
            "timer": "host wall clock between stream synchronisations, max over ranks"}


def _device_tables_to_host(be, torch, meta, dev_tables):
    """Download device-generated trace tables into pinned host arrays: the `host tables` legs prove from these."""
    pinned, host_pie = [], []
    for name, _ in meta:
        ptr, n_rows, n_cols = dev_tables[name]
        t = torch.empty((n_rows, n_cols), dtype=torch.int32, pin_memory=True)
        v = t.numpy().view(np.uint32)
        v[:] = be.download_ptr(ptr, n_rows * n_cols).reshape(n_rows, n_cols)
        pinned.append(t)
        host_pie.append((name, v))
    return host_pie, pinned


def bench_prove(be, torch, args):
    """luminair_prover::prover::prove on the cfg-3 graph (a + b over 2^log elements): device-resident tables (value), host
    tables in pinned memory (H2D of the trace + proof bytes back) and from the two input tensors (gen_trace on the device).
    Every table is produced by the product's own gen_trace (lb_trace_*); nothing here touches oracle/."""
    from luminair_b200.lookups import LookupLayout
    from luminair_b200.prover import STAGE_NAMES, last_stage_ms, prove
    from luminair_b200.trace import DeviceGraphTrace
    from luminair_b200.workloads import MLP_EXP2_RANGE, build_add_graph, build_mlp, build_wide, synthetic_add_graph_inputs
    log = args.prove_log
    a_raw, b_raw = synthetic_add_graph_inputs(log, seed=42)
    dg0 = build_add_graph(DeviceGraphTrace(be), a_raw, b_raw)
    meta0, dev, _ = dg0.finish()
    host_pie, pinned = _device_tables_to_host(be, torch, meta0, dev)
    reps = max(3, min(args.steps, 10))

    def run(device_tables):
        for _ in range(2):
            prove(host_pie, backend=be, device_tables=device_tables)
        ts, best_stages, nbytes = [], None, 0
        for _ in range(reps):
            t0 = time.perf_counter()
            proof = prove(host_pie, backend=be, device_tables=device_tables)
            dt = (time.perf_counter() - t0) * 1e3
            if not ts or dt < min(ts):
                best_stages = last_stage_ms(be)
            ts.append(dt)
            nbytes = len(proof)
        return ts, best_stages, nbytes

    ts_dev, st_dev, nbytes = run(dev)
    ts_host, st_host, _ = run(None)
    # the same proof from the two input TENSORS: upload a and b (pinned), gen_trace on the device (lb_trace_*), prove
    tens = []
    for x in (a_raw, b_raw):
        t = torch.empty(x.shape, dtype=torch.int32, pin_memory=True)
        t.numpy()[:] = x
        tens.append(t)

    def prove_from_tensors():
        dg = DeviceGraphTrace(be)
        dg.add(dg.input(tens[0].numpy()), dg.input(tens[1].numpy()))
        meta, dev_tables, _ = dg.finish()
        return prove(meta, backend=be, device_tables=dev_tables)

    ref_proof = prove(host_pie, backend=be)
    if prove_from_tensors() != ref_proof:
        raise SystemExit("bench: proof from device-generated tables differs from the proof from host tables")
    fixture = os.path.join(ROOT, "tests", "golden", "cfg3_add_log20.proof.bin")
    equals_fixture = None
    if log == 20 and os.path.exists(fixture):
        equals_fixture = open(fixture, "rb").read() == ref_proof
        if not equals_fixture:
            raise SystemExit("bench: the cfg-3 proof differs from the committed CPU-prover fixture")
    prove_from_tensors()
    ts_tens = []
    for _ in range(reps):
        t0 = time.perf_counter()
        prove_from_tensors()
        ts_tens.append((time.perf_counter() - t0) * 1e3)
    # the reference's one published prove() figure: Add 32x32 (docs/snippets/benchmark-component.mdx:173)
    sa, sb = synthetic_add_graph_inputs(10, seed=42)
    sm_rec = build_add_graph(DeviceGraphTrace(be), sa, sb)  # the recorder owns the device tables
    sm_meta, sm_dev, _ = sm_rec.finish()
    small, small_keep = _device_tables_to_host(be, torch, sm_meta, sm_dev)
    for _ in range(3):
        prove(small, backend=be)
    t0 = time.perf_counter()
    for _ in range(10):
        prove(small, backend=be)
    small_ms = (time.perf_counter() - t0) * 1e2
    h2d = sum(int(v.nbytes) for _, v in host_pie)
    # BASELINE configs[3] shape: the 2-64-64-1 tanh MLP of examples/black-schole-nn (synthetic weights), 7 components
    # including the Exp2 lookup table (2^17 rows) and the extended evaluation domain of its consumer.  The Exp2 layout is
    # the circuit setting a calibration run of this network yields (workloads.MLP_EXP2_RANGE)
    mlp_layouts = {"exp2": LookupLayout([MLP_EXP2_RANGE])}
    mlp_rec = build_mlp(DeviceGraphTrace(be))
    mlp_meta, mlp_dev_tables, _ = mlp_rec.finish(mlp_layouts)
    mlp_pre = mlp_rec.preprocessed
    mlp_pie, mlp_keep = _device_tables_to_host(be, torch, mlp_meta, mlp_dev_tables)
    for _ in range(2):
        prove(mlp_pie, backend=be, preprocessed=mlp_pre)
    t_mlp = []
    for _ in range(reps):
        t0 = time.perf_counter()
        mlp_proof = prove(mlp_pie, backend=be, preprocessed=mlp_pre)
        t_mlp.append((time.perf_counter() - t0) * 1e3)

    # the same proof with gen_trace on the device (lb_trace_op: Mul over broadcast operands, SumReduce, Add, Exp2 through the
    # host-generated LUT, Recip): only the weights / input / constants cross PCIe
    def mlp_from_tensors():
        dg = build_mlp(DeviceGraphTrace(be))
        meta, dev_tables, _ = dg.finish(mlp_layouts)
        return prove(meta, backend=be, device_tables=dev_tables, preprocessed=dg.preprocessed)

    if mlp_from_tensors() != mlp_proof:
        raise SystemExit("bench: MLP proof from device-generated tables differs from the proof from host tables")
    t_mlp_dev = []
    for _ in range(reps):
        t0 = time.perf_counter()
        mlp_from_tensors()
        t_mlp_dev.append((time.perf_counter() - t0) * 1e3)
    # compile once, run per execution (StwoCompiler once, gen_trace + prove per input): the recorded device graph is replayed
    # with a new network input; weights, gather indices, consumer counts and LUT columns stay resident
    from luminair_b200.lookups import to_fixed
    x_raw = to_fixed(np.asarray((15.0, 0.5), dtype=np.float64))

    def mlp_replay():
        mlp_rec.set_input(0, x_raw)
        meta, dev_tables, _ = mlp_rec.finish()
        return prove(meta, backend=be, device_tables=dev_tables, preprocessed=mlp_rec.preprocessed)

    if mlp_replay() != mlp_proof:
        raise SystemExit("bench: MLP proof from the replayed device graph differs from the proof from host tables")
    t_mlp_replay = []
    for _ in range(reps):
        t0 = time.perf_counter()
        mlp_replay()
        t_mlp_replay.append((time.perf_counter() - t0) * 1e3)
    mlp_info = {"workload": "prove(): Linear 2-64-64-1 with tanh (Mul/SumReduce/Add/Exp2+LUT/Recip/Inputs tables: "
                            + ", ".join(f"{k} {v.shape[0]}" for k, v in mlp_pie) + " rows), host tables (BASELINE configs[3] shape, "
                            "synthetic weights)",
                "ms_e2e_host_tables": {"min": min(t_mlp), "median": statistics.median(t_mlp)},
                "ms_e2e_host_tensors": {"min": min(t_mlp_dev), "median": statistics.median(t_mlp_dev),
                                        "how": "graph recording, LUT columns of the circuit settings (host libm as the reference; cached per layout), gen_trace "
                                               "on the device (lb_trace_op) and prove(); proof bytes identical to the host-table path"},
                "ms_e2e_recorded_graph": {"min": min(t_mlp_replay), "median": statistics.median(t_mlp_replay),
                                          "how": "graph recorded once (DeviceGraphTrace), per proof: upload the network input, replay "
                                                 "gen_trace on the device, prove()"},
                "proof_bytes": len(mlp_proof)}
    # BASELINE.json's headline trace shape: 2^log rows x 61 main-trace columns (Add + Mul + Rem + SumReduce tables over the same
    # two inputs) beside the Inputs table; device-resident tables
    wide_info = None
    if log >= 12:
        wrec = build_wide(DeviceGraphTrace(be), log)
        meta, wdev, wvalues = wrec.finish()
        shapes = {k: (wdev[k][1], wdev[k][2]) for k, _ in meta}
        for _ in range(2):
            prove(meta, backend=be, device_tables=wdev)
        t_w, st_w = [], None
        for _ in range(max(3, reps // 2)):
            t0 = time.perf_counter()
            wproof = prove(meta, backend=be, device_tables=wdev)
            dt = (time.perf_counter() - t0) * 1e3
            if not t_w or dt < min(t_w):
                st_w = last_stage_ms(be)
            t_w.append(dt)
        wfix = os.path.join(ROOT, "tests", "golden", "wide_log20.proof.bin")
        wide_equals_fixture = (open(wfix, "rb").read() == wproof) if (log == 20 and os.path.exists(wfix)) else None
        if wide_equals_fixture is False:
            raise SystemExit("bench: the wide proof differs from the committed CPU-prover fixture")
        # gen_trace of the same graph on the device from its two input tensors (pinned), then prove
        wtens = []
        for node in (0, 1):
            t = torch.empty((1 << log,), dtype=torch.int32, pin_memory=True)
            t.numpy().view(np.uint32)[:] = be.download_ptr(wvalues[node].ptr, 1 << log)
            wtens.append(t)

        def wide_from_tensors():
            dg = DeviceGraphTrace(be)
            a, b = dg.input(wtens[0].numpy()), dg.input(wtens[1].numpy())
            dg.add(a, b), dg.mul(a, b), dg.rem(a, b), dg.sum_reduce(a, 1)
            m, devt, _ = dg.finish()
            return prove(m, backend=be, device_tables=devt)

        if wide_from_tensors() != wproof:
            raise SystemExit("bench: wide proof from device-generated tables differs from the proof from the recorded graph's tables")
        t_wt = []
        for _ in range(max(3, reps // 2)):
            t0 = time.perf_counter()
            wide_from_tensors()
            t_wt.append((time.perf_counter() - t0) * 1e3)
        n_main_cols = sum(c for k, (r, c) in shapes.items() if k != "inputs")
        wide_info = {"workload": f"prove(): 2^{log} rows x {n_main_cols} main-trace columns (" +
                                 ", ".join(f"{k} {r}x{c}" for k, (r, c) in shapes.items()) + "), device-resident tables "
                                 "(the 2^20 x 64 trace shape of BASELINE.json's metric, built from real operator tables)",
                     "ms_device_resident": {"min": min(t_w), "median": statistics.median(t_w)},
                     "ms_e2e_host_tensors": {"min": min(t_wt), "median": statistics.median(t_wt), "h2d_bytes": int(2 * 4 << log),
                                             "how": "the two input tensors uploaded from pinned memory, all five tables "
                                                    "generated on the device (lb_trace_*), proof bytes identical"},
                     "proof_equals_cpu_prover_fixture": wide_equals_fixture,
                     "stages_ms": dict(zip(STAGE_NAMES, [round(x, 3) for x in st_w])), "proof_bytes": len(wproof)}
    return {
        "wide": wide_info,
        "mlp": mlp_info,
        "workload": f"prove(): a+b graph, Add 2^{log} rows x 15 cols + Inputs 2^{log + 1} rows x 7 cols, blow-up 2, Blake2s Merkle, FRI "
                    "(BASELINE configs[2]); proof bytes equal to the compiled CPU prover's at this size (tests/golden/cfg3_add_log20.proof.bin)",
        "proof_equals_cpu_prover_fixture": equals_fixture,
        "_proof_bytes": ref_proof,
        "ms_device_resident": {"min": min(ts_dev), "median": statistics.median(ts_dev)},
        "ms_e2e_host_tables": {"min": min(ts_host), "median": statistics.median(ts_host), "h2d_bytes": h2d, "d2h_bytes": nbytes},
        "ms_e2e_host_tensors": {"min": min(ts_tens), "median": statistics.median(ts_tens), "h2d_bytes": int(2 * 4 << log),
                                "d2h_bytes": nbytes, "how": "a and b uploaded from pinned memory, trace tables generated on the "
                                "device (lb_trace_inputs / lb_trace_add), proof bytes identical to the host-table path"},
        "proofs_per_s_e2e": 1e3 / statistics.median(ts_tens),
        "proofs_per_s_e2e_note": "gen_trace + prove from the input tensors in pinned host memory (host-table path: "
                                 f"{1e3 / statistics.median(ts_host):.1f} proofs/s)",
        "stages_ms_device_resident": dict(zip(STAGE_NAMES, [round(x, 3) for x in st_dev])),
        "proof_bytes": nbytes, "timer": "host wall clock around lb_prove (the call synchronises the stream before returning)",
        "add_32x32_prove_ms": small_ms,
        "published_reference_add_32x32_prove_ms": 13.05,
        "published_note": "GitHub Actions ubuntu-latest CPU, Criterion; other hardware, context only",
    }


def bench_concurrent_prove(args, expect_proof):
    """Throughput of K independent provers on ONE GPU: K contexts (own stream, own pool), one host thread each, every thread
    proving the same cfg-3 workload from its own device-resident tables.  A single proof is a dependent chain (Fiat-Shamir) with
    latency-bound stretches and ~12 host synchronisations; a second proof in flight fills them."""
    import threading
    from luminair_b200.backend import CudaBackend
    from luminair_b200.prover import prove
    from luminair_b200.trace import DeviceGraphTrace
    from luminair_b200.workloads import build_add_graph, synthetic_add_graph_inputs
    a_raw, b_raw = synthetic_add_graph_inputs(args.prove_log, seed=42)
    out = {}
    for k in (2, 3):
        bes = [CudaBackend(0) for _ in range(k)]
        try:
            recs = [build_add_graph(DeviceGraphTrace(b), a_raw, b_raw) for b in bes]
            work = [r.finish() for r in recs]
            reps = max(4, min(args.steps, 10))
            for b, (meta, dev, _) in zip(bes, work):
                if prove(meta, backend=b, device_tables=dev) != expect_proof:
                    raise SystemExit("bench: a concurrent context produced different proof bytes")
            gate = threading.Barrier(k + 1)
            bad = []

            def worker(b, meta, dev):
                gate.wait()
                for _ in range(reps):
                    if len(prove(meta, backend=b, device_tables=dev)) != len(expect_proof):
                        bad.append(1)
                gate.wait()

            th = [threading.Thread(target=worker, args=(b, w[0], w[1])) for b, w in zip(bes, work)]
            for t in th:
                t.start()
            gate.wait()
            t0 = time.perf_counter()
            gate.wait()
            dt = time.perf_counter() - t0
            for t in th:
                t.join()
            if bad:
                raise SystemExit("bench: concurrent proofs differ in size")
            out[f"contexts_{k}"] = {"proofs_per_s": k * reps / dt, "ms_per_proof_amortised": dt * 1e3 / (k * reps), "proofs": k * reps}
        finally:
            for b in bes:
                b.close()
    out["how"] = ("K contexts on cuda:0, one host thread each (ctypes releases the GIL inside lb_prove), device-resident tables, "
                  "wall clock from a common start barrier to the last proof; every context's proof checked against the "
                  "single-context bytes first")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-prove", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--sharded-log", type=int, default=22)
    ap.add_argument("--sharded-cols-per-gpu", type=int, default=32)
    ap.add_argument("--prove-log", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
