"""CPU-side checks of the drop-in boundary: the shared library loads and exports every
symbol include/luminair_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "luminair_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from luminair_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # the ctypes mirror binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from luminair_b200._lib import LuminairB200Error
    from luminair_b200.backend import CudaBackend
    with pytest.raises(LuminairB200Error, match="no CPU fallback"):
        CudaBackend(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "luminair_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_rust_ffi_file_is_in_sync_with_the_header():
    """integration/rust/luminair_b200_sys.rs is generated from the header (scripts/gen_rust_ffi.py): it is up to date and
    declares every function the library exports."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_rust_ffi.py"), "--check"]).returncode == 0
    rs = open(os.path.join(root, "integration", "rust", "luminair_b200_sys.rs")).read()
    declared = set(re.findall(r"pub fn (lb_\w+)\(", rs))
    from luminair_b200._lib import SIGNATURES
    assert declared == set(SIGNATURES)
