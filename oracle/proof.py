"""LuminairProof <-> bincode 1.3 bytes (oracle; test infrastructure only).

Layout restated from crates/prover/src/lib.rs:15-32 (``LuminairProof{claim,
interaction_claim, proof}`` + ``bincode::serialize``), crates/air/src/lib.rs:29-48,
189-207 (N Option fields) and stwo's ``CommitmentSchemeProof`` / ``FriProof``
serde derives; checked byte-for-byte (parse -> serialise round trip to EOF)
against the committed artifact ui/demo/public/proof.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional

from .fields import QM31

N_COMPONENTS = 17  # crates/air/src/lib.rs:30-48 ; the UI artifact predates this and has 8


@dataclass
class MerkleDecommitment:
    hash_witness: List[bytes] = field(default_factory=list)
    column_witness: List[int] = field(default_factory=list)


@dataclass
class FriLayerProof:
    fri_witness: List[QM31]
    decommitment: MerkleDecommitment
    commitment: bytes


@dataclass
class FriProof:
    first_layer: FriLayerProof
    inner_layers: List[FriLayerProof]
    last_layer_poly: List[QM31]
    last_layer_log_size: int


@dataclass
class PcsConfig:
    pow_bits: int = 5
    log_blowup_factor: int = 1
    log_last_layer_degree_bound: int = 0
    n_queries: int = 3


@dataclass
class StarkProof:
    config: PcsConfig
    commitments: List[bytes]
    sampled_values: List[List[List[QM31]]]  # tree -> column -> samples
    decommitments: List[MerkleDecommitment]
    queried_values: List[List[int]]  # tree -> flat
    proof_of_work: int
    fri_proof: FriProof


@dataclass
class LuminairProof:
    claim: List[Optional[int]]  # per component slot: log_size or None
    interaction_claim: List[Optional[QM31]]  # per slot: claimed_sum or None
    proof: StarkProof


class _R:
    def __init__(self, b: bytes):
        self.b = b
        self.o = 0

    def u8(self):
        v = self.b[self.o]
        self.o += 1
        return v

    def u32(self):
        v = struct.unpack_from("<I", self.b, self.o)[0]
        self.o += 4
        return v

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def raw(self, n):
        v = self.b[self.o : self.o + n]
        assert len(v) == n, "unexpected EOF"
        self.o += n
        return v

    def qm31(self):
        return QM31(*struct.unpack("<4I", self.raw(16)))


def _read_decommitment(r: _R) -> MerkleDecommitment:
    hw = [r.raw(32) for _ in range(r.u64())]
    cw = [r.u32() for _ in range(r.u64())]
    return MerkleDecommitment(hw, cw)


def _read_fri_layer(r: _R) -> FriLayerProof:
    w = [r.qm31() for _ in range(r.u64())]
    d = _read_decommitment(r)
    c = r.raw(32)
    return FriLayerProof(w, d, c)


def from_bincode(data: bytes, n_components: int = N_COMPONENTS) -> LuminairProof:
    r = _R(data)
    claim = []
    for _ in range(n_components):
        tag = r.u8()
        assert tag in (0, 1)
        claim.append(r.u32() if tag else None)
    iclaim = []
    for _ in range(n_components):
        tag = r.u8()
        assert tag in (0, 1)
        iclaim.append(r.qm31() if tag else None)
    cfg = PcsConfig(r.u32(), r.u32(), r.u32(), r.u64())
    commitments = [r.raw(32) for _ in range(r.u64())]
    sampled = []
    for _ in range(r.u64()):
        cols = []
        for _ in range(r.u64()):
            cols.append([r.qm31() for _ in range(r.u64())])
        sampled.append(cols)
    decommitments = [_read_decommitment(r) for _ in range(r.u64())]
    queried = []
    for _ in range(r.u64()):
        queried.append([r.u32() for _ in range(r.u64())])
    pow_nonce = r.u64()
    first = _read_fri_layer(r)
    inner = [_read_fri_layer(r) for _ in range(r.u64())]
    coeffs = [r.qm31() for _ in range(r.u64())]
    log_size = r.u32()
    assert r.o == len(data), f"trailing bytes: {len(data) - r.o}"
    return LuminairProof(
        claim, iclaim,
        StarkProof(cfg, commitments, sampled, decommitments, queried, pow_nonce,
                   FriProof(first, inner, coeffs, log_size)),
    )


def _w_qm31(out, q: QM31):
    out.append(struct.pack("<4I", *q.tup()))


def _w_decommitment(out, d: MerkleDecommitment):
    out.append(struct.pack("<Q", len(d.hash_witness)))
    out.extend(d.hash_witness)
    out.append(struct.pack("<Q", len(d.column_witness)))
    out.append(struct.pack(f"<{len(d.column_witness)}I", *d.column_witness))


def _w_fri_layer(out, l: FriLayerProof):
    out.append(struct.pack("<Q", len(l.fri_witness)))
    for q in l.fri_witness:
        _w_qm31(out, q)
    _w_decommitment(out, l.decommitment)
    out.append(l.commitment)


def to_bincode(p: LuminairProof) -> bytes:
    out = []
    for c in p.claim:
        out.append(b"\0" if c is None else b"\x01" + struct.pack("<I", c))
    for c in p.interaction_claim:
        if c is None:
            out.append(b"\0")
        else:
            out.append(b"\x01")
            _w_qm31(out, c)
    s = p.proof
    out.append(struct.pack("<IIIQ", s.config.pow_bits, s.config.log_blowup_factor,
                           s.config.log_last_layer_degree_bound, s.config.n_queries))
    out.append(struct.pack("<Q", len(s.commitments)))
    out.extend(s.commitments)
    out.append(struct.pack("<Q", len(s.sampled_values)))
    for tree in s.sampled_values:
        out.append(struct.pack("<Q", len(tree)))
        for col in tree:
            out.append(struct.pack("<Q", len(col)))
            for q in col:
                _w_qm31(out, q)
    out.append(struct.pack("<Q", len(s.decommitments)))
    for d in s.decommitments:
        _w_decommitment(out, d)
    out.append(struct.pack("<Q", len(s.queried_values)))
    for qv in s.queried_values:
        out.append(struct.pack("<Q", len(qv)))
        out.append(struct.pack(f"<{len(qv)}I", *qv))
    out.append(struct.pack("<Q", s.proof_of_work))
    f = s.fri_proof
    _w_fri_layer(out, f.first_layer)
    out.append(struct.pack("<Q", len(f.inner_layers)))
    for l in f.inner_layers:
        _w_fri_layer(out, l)
    out.append(struct.pack("<Q", len(f.last_layer_poly)))
    for q in f.last_layer_poly:
        _w_qm31(out, q)
    out.append(struct.pack("<I", f.last_layer_log_size))
    return b"".join(out)
