"""``LuminairProof`` wire formats: bincode 1.3 and JSON, as ``to_bincode / from_bincode / to_json / from_json`` of
/root/reference/crates/prover/src/lib.rs:22-107 produce them.  Host-side data format next to the path - no compute.

``lb_prove`` emits the bincode form; this module parses it into a plain structure and re-serialises it either way, so a proof can
be handed to a consumer of the reference's JSON files (``to_json_file``, lib.rs:80-96) and read back.

Field names follow the serde derives: ``LuminairProof {claim, interaction_claim, proof}`` (lib.rs:15-20), ``LuminairClaim`` /
``LuminairInteractionClaim`` with one ``Option`` per component (crates/air/src/lib.rs:30-48, 189-207), ``Claim {log_size,
_marker}`` (``PhantomData`` serialises as a unit: nothing in bincode, ``null`` in JSON; components/mod.rs:148-152),
``InteractionClaim {claimed_sum}`` (components/mod.rs:202-205).  ``StarkProof`` and what it contains are stwo's (un-vendored,
rev 0790eba): ``StarkProof(CommitmentSchemeProof)`` newtype, ``CommitmentSchemeProof {config {pow_bits, fri_config
{log_blowup_factor, log_last_layer_degree_bound, n_queries}}, commitments, sampled_values, decommitments {hash_witness,
column_witness}, queried_values, proof_of_work, fri_proof {first_layer {fri_witness, decommitment, commitment}, inner_layers,
last_layer_poly {coeffs, log_size}}}``; ``M31(u32)`` is a number, ``CM31(M31, M31)`` / ``QM31(CM31, CM31)`` nested pairs,
``Blake2sHash([u8; 32])`` an array of 32 numbers.  The bincode layout is pinned by the reference's committed proof
(tests/golden/demo_proof.bin round-trips to the byte); the JSON names of the stwo part are restated from recall - parity
unpinned (the reference commits no JSON proof).
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

COMPONENTS = ["add", "mul", "recip", "sin", "sin_lookup", "sum_reduce", "max_reduce", "sqrt", "rem", "exp2", "exp2_lookup",
              "log2", "log2_lookup", "less_than", "range_check_lookup", "inputs", "contiguous"]  # crates/air/src/lib.rs:30-48
Felt = Tuple[int, int, int, int]  # QM31 coordinates (a.re, a.im, b.re, b.im)


@dataclass
class Decommitment:
    hash_witness: List[bytes] = field(default_factory=list)
    column_witness: List[int] = field(default_factory=list)


@dataclass
class FriLayer:
    fri_witness: List[Felt]
    decommitment: Decommitment
    commitment: bytes


@dataclass
class Proof:
    claim: List[Optional[int]]
    interaction_claim: List[Optional[Felt]]
    pow_bits: int
    log_blowup_factor: int
    log_last_layer_degree_bound: int
    n_queries: int
    commitments: List[bytes]
    sampled_values: List[List[List[Felt]]]
    decommitments: List[Decommitment]
    queried_values: List[List[int]]
    proof_of_work: int
    first_layer: FriLayer
    inner_layers: List[FriLayer]
    last_layer_coeffs: List[Felt]
    last_layer_log_size: int


class _Reader:
    def __init__(self, data: bytes):
        self.b, self.o = data, 0

    def take(self, fmt):
        if self.o + struct.calcsize(fmt) > len(self.b):
            raise ValueError("unexpected end of proof")
        v = struct.unpack_from(fmt, self.b, self.o)
        self.o += struct.calcsize(fmt)
        return v if len(v) > 1 else v[0]

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        if len(v) != n:
            raise ValueError("unexpected end of proof")
        self.o += n
        return v

    def felt(self) -> Felt:
        return tuple(self.take("<4I"))

    def decommitment(self) -> Decommitment:
        hw = [self.raw(32) for _ in range(self.take("<Q"))]
        n = self.take("<Q")
        return Decommitment(hw, list(struct.unpack_from(f"<{n}I", self.raw(4 * n))) if n else [])

    def fri_layer(self) -> FriLayer:
        w = [self.felt() for _ in range(self.take("<Q"))]
        return FriLayer(w, self.decommitment(), self.raw(32))


def from_bincode(data: bytes, n_components: int = len(COMPONENTS)) -> Proof:
    r = _Reader(data)

    def option(read):
        tag = r.take("<B")
        if tag > 1:
            raise ValueError("bad Option tag")
        return read() if tag else None

    claim = [option(lambda: r.take("<I")) for _ in range(n_components)]
    iclaim = [option(r.felt) for _ in range(n_components)]
    pow_bits, blow, last = r.take("<III")
    n_queries = r.take("<Q")
    commitments = [r.raw(32) for _ in range(r.take("<Q"))]
    sampled = [[[r.felt() for _ in range(r.take("<Q"))] for _ in range(r.take("<Q"))] for _ in range(r.take("<Q"))]
    decommitments = [r.decommitment() for _ in range(r.take("<Q"))]
    queried = []
    for _ in range(r.take("<Q")):
        n = r.take("<Q")
        queried.append(list(struct.unpack_from(f"<{n}I", r.raw(4 * n))) if n else [])
    nonce = r.take("<Q")
    first = r.fri_layer()
    inner = [r.fri_layer() for _ in range(r.take("<Q"))]
    coeffs = [r.felt() for _ in range(r.take("<Q"))]
    log_size = r.take("<I")
    if r.o != len(data):
        raise ValueError(f"trailing bytes after the proof: {len(data) - r.o}")
    return Proof(claim, iclaim, pow_bits, blow, last, n_queries, commitments, sampled, decommitments, queried, nonce, first, inner,
                 coeffs, log_size)


def to_bincode(p: Proof) -> bytes:
    out = bytearray()

    def felt(f):
        out.extend(struct.pack("<4I", *f))

    def decommitment(d):
        out.extend(struct.pack("<Q", len(d.hash_witness)))
        for h in d.hash_witness:
            out.extend(h)
        out.extend(struct.pack("<Q", len(d.column_witness)))
        out.extend(struct.pack(f"<{len(d.column_witness)}I", *d.column_witness))

    def fri_layer(l):
        out.extend(struct.pack("<Q", len(l.fri_witness)))
        for f in l.fri_witness:
            felt(f)
        decommitment(l.decommitment)
        out.extend(l.commitment)

    for c in p.claim:
        out.extend(b"\0" if c is None else b"\x01" + struct.pack("<I", c))
    for c in p.interaction_claim:
        if c is None:
            out.extend(b"\0")
        else:
            out.extend(b"\x01")
            felt(c)
    out.extend(struct.pack("<IIIQ", p.pow_bits, p.log_blowup_factor, p.log_last_layer_degree_bound, p.n_queries))
    out.extend(struct.pack("<Q", len(p.commitments)))
    for h in p.commitments:
        out.extend(h)
    out.extend(struct.pack("<Q", len(p.sampled_values)))
    for tree in p.sampled_values:
        out.extend(struct.pack("<Q", len(tree)))
        for col in tree:
            out.extend(struct.pack("<Q", len(col)))
            for f in col:
                felt(f)
    out.extend(struct.pack("<Q", len(p.decommitments)))
    for d in p.decommitments:
        decommitment(d)
    out.extend(struct.pack("<Q", len(p.queried_values)))
    for qv in p.queried_values:
        out.extend(struct.pack("<Q", len(qv)))
        out.extend(struct.pack(f"<{len(qv)}I", *qv))
    out.extend(struct.pack("<Q", p.proof_of_work))
    fri_layer(p.first_layer)
    out.extend(struct.pack("<Q", len(p.inner_layers)))
    for l in p.inner_layers:
        fri_layer(l)
    out.extend(struct.pack("<Q", len(p.last_layer_coeffs)))
    for f in p.last_layer_coeffs:
        felt(f)
    out.extend(struct.pack("<I", p.last_layer_log_size))
    return bytes(out)


# ---- JSON (serde_json of the same structs) ---------------------------------------------------------------------------
def _felt_json(f: Felt):
    return [[f[0], f[1]], [f[2], f[3]]]


def _felt_from(j) -> Felt:
    (a, b), (c, d) = j
    return (int(a), int(b), int(c), int(d))


def _hash_json(h: bytes):
    return list(h)


def _decommitment_json(d: Decommitment):
    return {"hash_witness": [_hash_json(h) for h in d.hash_witness], "column_witness": list(d.column_witness)}


def _decommitment_from(j) -> Decommitment:
    return Decommitment([bytes(h) for h in j["hash_witness"]], [int(x) for x in j["column_witness"]])


def _fri_layer_json(l: FriLayer):
    return {"fri_witness": [_felt_json(f) for f in l.fri_witness], "decommitment": _decommitment_json(l.decommitment),
            "commitment": _hash_json(l.commitment)}


def _fri_layer_from(j) -> FriLayer:
    return FriLayer([_felt_from(f) for f in j["fri_witness"]], _decommitment_from(j["decommitment"]), bytes(j["commitment"]))


def to_json(p: Proof, components=COMPONENTS) -> str:
    """``LuminairProof::to_json`` (serde_json::to_string_pretty, lib.rs:64-68)."""
    if len(p.claim) != len(components):
        raise ValueError("claim slots do not match the component list")
    doc = {
        "claim": {name: (None if c is None else {"log_size": c, "_marker": None}) for name, c in zip(components, p.claim)},
        "interaction_claim": {name: (None if c is None else {"claimed_sum": _felt_json(c)})
                              for name, c in zip(components, p.interaction_claim)},
        "proof": {
            "config": {"pow_bits": p.pow_bits,
                       "fri_config": {"log_blowup_factor": p.log_blowup_factor,
                                      "log_last_layer_degree_bound": p.log_last_layer_degree_bound, "n_queries": p.n_queries}},
            "commitments": [_hash_json(h) for h in p.commitments],
            "sampled_values": [[[_felt_json(f) for f in col] for col in tree] for tree in p.sampled_values],
            "decommitments": [_decommitment_json(d) for d in p.decommitments],
            "queried_values": [list(qv) for qv in p.queried_values],
            "proof_of_work": p.proof_of_work,
            "fri_proof": {"first_layer": _fri_layer_json(p.first_layer),
                          "inner_layers": [_fri_layer_json(l) for l in p.inner_layers],
                          "last_layer_poly": {"coeffs": [_felt_json(f) for f in p.last_layer_coeffs],
                                              "log_size": p.last_layer_log_size}},
        },
    }
    return json.dumps(doc, indent=2)


def from_json(text: str, components=COMPONENTS) -> Proof:
    """``LuminairProof::from_json`` (lib.rs:71-78)."""
    doc = json.loads(text)
    claim = [None if doc["claim"].get(n) is None else int(doc["claim"][n]["log_size"]) for n in components]
    iclaim = [None if doc["interaction_claim"].get(n) is None else _felt_from(doc["interaction_claim"][n]["claimed_sum"])
              for n in components]
    pr = doc["proof"]
    cfg, fri = pr["config"], pr["fri_proof"]
    return Proof(claim, iclaim, int(cfg["pow_bits"]), int(cfg["fri_config"]["log_blowup_factor"]),
                 int(cfg["fri_config"]["log_last_layer_degree_bound"]), int(cfg["fri_config"]["n_queries"]),
                 [bytes(h) for h in pr["commitments"]],
                 [[[_felt_from(f) for f in col] for col in tree] for tree in pr["sampled_values"]],
                 [_decommitment_from(d) for d in pr["decommitments"]],
                 [[int(x) for x in qv] for qv in pr["queried_values"]], int(pr["proof_of_work"]),
                 _fri_layer_from(fri["first_layer"]), [_fri_layer_from(l) for l in fri["inner_layers"]],
                 [_felt_from(f) for f in fri["last_layer_poly"]["coeffs"]], int(fri["last_layer_poly"]["log_size"]))


def bincode_to_json(data: bytes) -> str:
    return to_json(from_bincode(data))


def json_to_bincode(text: str) -> bytes:
    return to_bincode(from_json(text))


def write_json_file(path: str, data: bytes):
    """``to_json_file`` (lib.rs:80-96) for a bincode proof as ``lb_prove`` returns it."""
    with open(path, "w") as f:
        f.write(bincode_to_json(data))


def read_json_file(path: str) -> bytes:
    """``from_json_file`` (lib.rs:98-107) -> the bincode form."""
    with open(path) as f:
        return json_to_bincode(f.read())
