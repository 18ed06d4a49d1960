"""luminair_b200 — B200-native Circle-STARK prover backend behind LuminAIR's prove() path.

The compute path is ``libluminair_b200.so`` (hand-written sm_100a CUDA, C ABI in
``include/luminair_b200.h``).  This package is the thin host mirror of the reference
interface (stwo backend traits + ``luminair_prover::prover::prove``); it never falls
back to a CPU implementation and raises if the CUDA library or a GPU is missing.
"""
from ._lib import LuminairB200Error, load_library  # noqa: F401

__all__ = ["LuminairB200Error", "load_library"]
