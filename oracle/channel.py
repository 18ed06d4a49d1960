"""Blake2s Fiat-Shamir channel (oracle; test infrastructure only).

Restates stwo ``core/channel/blake2s.rs`` (reference call sites:
crates/prover/src/prover.rs:44,177,296; crates/air/src/components/mod.rs:173-176,
209-211,227-235).  Two variants are switchable (SURVEY.md Appendix A hazards):

  "legacy"  mix_u64 = raw compress(h=digest, m=[lo,hi,0..], t=0, f=0); draw =
            Blake2s(digest || counter as 32 LE bytes).  ARTIFACT-VERIFIED against
            ui/demo/public/proof (tests/test_oracle_kat.py).
  "v2"      mix_u64 = Blake2s(digest || lo || hi); draw appends one 0x00 byte.
            [stwo-recall; unpinned]
"""
from __future__ import annotations

import struct

from . import blake2s
from .fields import P, QM31

FELTS_PER_HASH = 8


class Blake2sChannel:
    def __init__(self, variant: str = "legacy"):
        assert variant in ("legacy", "v2")
        self.variant = variant
        self.digest = bytes(32)
        self.n_challenges = 0
        self.n_sent = 0
        self.log = []  # (kind, payload) trace of mixes, for transcript tests

    def _update(self, d: bytes, kind, payload=None):
        self.digest = d
        self.n_challenges += 1
        self.n_sent = 0
        self.log.append((kind, payload))

    def mix_root(self, root: bytes):
        """Blake2sMerkleChannel::mix_root: Blake2s(digest || root)."""
        self._update(blake2s.hash(self.digest + root), "root", root)

    def mix_felts(self, felts):
        data = b"".join(struct.pack("<4I", *f.tup()) for f in felts)
        self._update(blake2s.hash(self.digest + data), "felts", len(felts))

    def mix_u32s(self, words):
        data = struct.pack(f"<{len(words)}I", *words)
        self._update(blake2s.hash(self.digest + data), "u32s", list(words))

    def mix_u64(self, v: int):
        lo, hi = v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF
        if self.variant == "legacy":
            h = struct.unpack("<8I", self.digest)
            res = blake2s.compress(h, [lo, hi] + [0] * 14, 0, 0, 0, 0)
            self._update(struct.pack("<8I", *res), "u64", v)
        else:
            data = struct.pack("<2I", lo, hi)
            self._update(blake2s.hash(self.digest + data), "u64", v)

    def trailing_zeros(self) -> int:
        v = int.from_bytes(self.digest[:16], "little")
        if v == 0:
            return 128
        return (v & -v).bit_length() - 1

    def draw_random_bytes(self) -> bytes:
        inp = self.digest + self.n_sent.to_bytes(32, "little")
        if self.variant == "v2":
            inp += b"\0"
        self.n_sent += 1
        return blake2s.hash(inp)

    def draw_base_felts(self):
        while True:
            u = struct.unpack("<8I", self.draw_random_bytes())
            if all(x < 2 * P for x in u):
                return [x % P for x in u]

    def draw_secure_felt(self) -> QM31:
        f = self.draw_base_felts()
        return QM31(*f[:4])

    def draw_secure_felts(self, n: int):
        out = []
        pool = []
        while len(out) < n:
            if len(pool) < 4:
                pool += self.draw_base_felts()
            out.append(QM31(*pool[:4]))
            pool = pool[4:]
        return out


def grind(channel: Blake2sChannel, pow_bits: int) -> int:
    """GrindOps<Blake2sChannel>::grind: smallest nonce such that after
    ``mix_u64(nonce)`` the digest has >= pow_bits trailing zeros."""
    nonce = 0
    if channel.variant == "legacy":
        h = struct.unpack("<8I", channel.digest)
        mask = (1 << pow_bits) - 1
        assert pow_bits <= 32
        while True:
            res = blake2s.compress(h, [nonce & 0xFFFFFFFF, nonce >> 32] + [0] * 14, 0, 0, 0, 0)
            if res[0] & mask == 0:
                return nonce
            nonce += 1
    while True:
        d = blake2s.hash(channel.digest + struct.pack("<2I", nonce & 0xFFFFFFFF, nonce >> 32))
        v = int.from_bytes(d[:16], "little")
        tz = 128 if v == 0 else (v & -v).bit_length() - 1
        if tz >= pow_bits:
            return nonce
        nonce += 1
