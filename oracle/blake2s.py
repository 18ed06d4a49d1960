"""Blake2s-256 helpers (oracle; test infrastructure only).

``hash`` is hashlib's standard Blake2s-256; ``compress`` is the raw RFC 7693
compression function F(h, m, t, f) that stwo's legacy ``Blake2sChannel::mix_u64``
and ``GrindOps`` use directly (SURVEY.md Appendix A "version hazards").
"""
from __future__ import annotations

import hashlib
import struct

IV = (
    0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A,
    0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19,
)
SIGMA = (
    (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15),
    (14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3),
    (11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4),
    (7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8),
    (9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13),
    (2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9),
    (12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11),
    (13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10),
    (6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5),
    (10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0),
)
M32 = 0xFFFFFFFF


def _rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & M32


def compress(h, m, t0=0, t1=0, f0=0, f1=0):
    """RFC 7693 F: h[8], m[16] u32 words -> h'[8]."""
    v = list(h) + list(IV)
    v[12] ^= t0
    v[13] ^= t1
    v[14] ^= f0
    v[15] ^= f1

    def g(a, b, c, d, x, y):
        v[a] = (v[a] + v[b] + x) & M32
        v[d] = _rotr(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & M32
        v[b] = _rotr(v[b] ^ v[c], 12)
        v[a] = (v[a] + v[b] + y) & M32
        v[d] = _rotr(v[d] ^ v[a], 8)
        v[c] = (v[c] + v[d]) & M32
        v[b] = _rotr(v[b] ^ v[c], 7)

    for r in range(10):
        s = SIGMA[r]
        g(0, 4, 8, 12, m[s[0]], m[s[1]])
        g(1, 5, 9, 13, m[s[2]], m[s[3]])
        g(2, 6, 10, 14, m[s[4]], m[s[5]])
        g(3, 7, 11, 15, m[s[6]], m[s[7]])
        g(0, 5, 10, 15, m[s[8]], m[s[9]])
        g(1, 6, 11, 12, m[s[10]], m[s[11]])
        g(2, 7, 8, 13, m[s[12]], m[s[13]])
        g(3, 4, 9, 14, m[s[14]], m[s[15]])
    return [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]


def hash(data: bytes) -> bytes:  # noqa: A001
    return hashlib.blake2s(data, digest_size=32).digest()


def hash_pure(data: bytes) -> bytes:
    """Blake2s-256 built on `compress` (checks `compress` against hashlib)."""
    h = list(IV)
    h[0] ^= 0x01010020
    n = len(data)
    blocks = [data[i : i + 64] for i in range(0, max(n, 1), 64)] or [b""]
    t = 0
    for i, blk in enumerate(blocks):
        last = i == len(blocks) - 1
        t += len(blk)
        blk = blk.ljust(64, b"\0")
        m = struct.unpack("<16I", blk)
        h = compress(h, m, t & M32, t >> 32, M32 if last else 0, 0)
    return struct.pack("<8I", *h)
