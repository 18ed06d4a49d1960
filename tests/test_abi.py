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


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct the ctypes mirror passes by pointer, computed by gcc from the header itself, and the
    LB_OP_* / LB_REL_* constants the Python side hard-codes."""
    import ctypes as C
    import subprocess
    from luminair_b200 import _lib, trace
    structs = {"lb_trace_table": _lib.TraceTable, "lb_prove_config": _lib.ProveConfig,
               "lb_preprocessed_column": _lib.PreprocessedColumn, "lb_relation": _lib.Relation,
               "lb_batch_shard": _lib.BatchShard, "lb_sample_batch": _lib.SampleBatch, "lb_lookup": _lib.Lookup,
               "lb_trace_op_desc": _lib.TraceOpDesc}
    ops = {"add": "LB_OP_ADD", "mul": "LB_OP_MUL", "recip": "LB_OP_RECIP", "sin": "LB_OP_SIN", "sum_reduce": "LB_OP_SUM_REDUCE",
           "max_reduce": "LB_OP_MAX_REDUCE", "sqrt": "LB_OP_SQRT", "rem": "LB_OP_REM", "exp2": "LB_OP_EXP2", "log2": "LB_OP_LOG2",
           "less_than": "LB_OP_LESS_THAN", "inputs": "LB_OP_INPUTS", "contiguous": "LB_OP_CONTIGUOUS"}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "luminair_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    for key, macro in ops.items():
        lines.append(f'printf("op.{key} %d\\n", {macro});')
    lines.append('return 0; }')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    for key in ops:
        assert int(got[f"op.{key}"]) == trace.OP_CODE[key], key
    # the operator code is the claim slot of the operator's component
    from luminair_b200.prover import CLAIM_SLOT
    assert all(trace.OP_CODE[k] == CLAIM_SLOT[k] for k in trace.OP_CODE)


def test_generate_secure_powers_on_the_host():
    """AccumulationOps::generate_secure_powers needs no device: [1, f, f^2, ...] against the CPU restatement's QM31."""
    import ctypes as C
    import numpy as np
    from luminair_b200._lib import load_library
    from oracle.fields import QM31
    lib = load_library()
    felt = (123456789, 987654321, 5, 2147483646)
    out = np.zeros((9, 4), dtype=np.uint32)
    rc = lib.lb_generate_secure_powers((C.c_uint32 * 4)(*felt), 9, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == 0
    acc, f = QM31(1, 0, 0, 0), QM31(*felt)
    for k in range(9):
        assert tuple(int(x) for x in out[k]) == acc.tup()
        acc = acc * f
    assert lib.lb_generate_secure_powers((C.c_uint32 * 4)(2147483647, 0, 0, 0), 1, out.ctypes.data_as(C.POINTER(C.c_uint32))) == -3
