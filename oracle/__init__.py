"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain numpy / pure-Python restatement of the Circle-STARK algorithms that
LuminAIR reaches through the un-vendored crate ``stwo`` (git rev ``0790eba``,
/root/reference/Cargo.toml:21-28).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; the product (``luminair_b200``) never does and fails loudly when its
CUDA library is missing.

Parity pin: ``oracle.kat`` runs the restated *verifier* over the one proof
artifact the reference commits (``ui/demo/public/proof``; a copy of its bytes
lives in ``tests/golden/demo_proof.bin``).  See DESIGN.md "Oracle".
"""
