"""Throughput of K prover contexts sharing one GPU (bench.py's `prove.concurrent` leg on its own):
    python scripts/concurrent_prove.py [--log 20] [--reps 10]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log", type=int, default=20)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    import bench
    from luminair_b200.backend import CudaBackend
    from luminair_b200.prover import prove
    from luminair_b200.trace import DeviceGraphTrace
    from luminair_b200.workloads import build_add_graph, synthetic_add_graph_inputs
    be = CudaBackend(0)
    rec = build_add_graph(DeviceGraphTrace(be), *synthetic_add_graph_inputs(a.log, seed=42))
    meta, dev, _ = rec.finish()
    proof = prove(meta, backend=be, device_tables=dev)
    ns = argparse.Namespace(prove_log=a.log, steps=a.reps)
    print(json.dumps(bench.bench_concurrent_prove(ns, proof)))
    be.close()


if __name__ == "__main__":
    main()
