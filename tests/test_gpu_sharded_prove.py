"""lb_prove_sharded on >= 2 GPUs (one process per GPU under torch.distributed.run): the proof bytes equal the single-GPU
lb_prove of the same tables on every rank.  Skipped on boxes with one GPU (the driver's round-end GPU test box); run by
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded_prove.py -m gpu` and recorded under profiles/."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_proof_equals_single_gpu_proof(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "scripts", "run_sharded_prove.py"), "14", "--all-components"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world
    for name in ("cfg3_add", "wide"):
        assert out[name]["proof_equals_single_device"] is True
        assert out[name]["nccl"]["bytes_sent"] > 0
    assert out["all_components"]["proof_equals_single_device"] is True  # 17 components, LUT tree, two PcsConfigs / channels


def test_comm_entry_points_fail_cleanly_without_a_gpu():
    """No GPU here: the communicator cannot be built, and the failure is a return code, not a crash."""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from luminair_b200._lib import load_library
    lib = load_library()
    out = C.c_void_p()
    assert lib.lb_comm_init(None, b"\0" * 128, 0, 2, C.byref(out)) == -3
    assert lib.lb_comm_stats(None, None, None, None, None, None) == -3
    lib.lb_comm_destroy(None)
