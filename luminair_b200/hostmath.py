"""Scalar host-side arithmetic of the transcript: M31 / QM31 values, circle-group points, the Blake2s channel.

This is what stays on the host in the reference as well (stwo ``Blake2sChannel``, ``CirclePoint<SecureField>``,
``SecureField`` scalars inside ``stwo::prover::prove``; reached from /root/reference/crates/prover/src/prover.rs:44,177,296,
311-312): a few dozen field operations and hashes per proof, never a loop over a column.  ``luminair_b200.traits_prover`` uses
it to drive the backend-trait entry points of the C ABI; the compiled host (csrc/prover.cu) has its own C++ copy.
"""
from __future__ import annotations

import hashlib
import struct

P = (1 << 31) - 1


def m_inv(a: int) -> int:
    return pow(a, P - 2, P)


class QM31:
    """(a0 + a1 i) + (b0 + b1 i) u with i^2 = -1, u^2 = 2 + i; coordinates are ints in [0, P)."""
    __slots__ = ("c",)

    def __init__(self, c0=0, c1=0, c2=0, c3=0):
        self.c = (c0 % P, c1 % P, c2 % P, c3 % P)

    def __add__(self, o):
        o = _q(o)
        return QM31(*[(x + y) for x, y in zip(self.c, o.c)])

    def __sub__(self, o):
        o = _q(o)
        return QM31(*[(x - y) for x, y in zip(self.c, o.c)])

    def __neg__(self):
        return QM31(*[-x for x in self.c])

    def __mul__(self, o):
        if isinstance(o, int):
            return QM31(*[x * o for x in self.c])
        a0, a1, b0, b1 = self.c
        c0, c1, d0, d1 = o.c
        # (a + b u)(c + d u) = ac + (2 + i) bd + (ad + bc) u over CM31
        ac = (a0 * c0 - a1 * c1, a0 * c1 + a1 * c0)
        bd = (b0 * d0 - b1 * d1, b0 * d1 + b1 * d0)
        ad = (a0 * d0 - a1 * d1, a0 * d1 + a1 * d0)
        bc = (b0 * c0 - b1 * c1, b0 * c1 + b1 * c0)
        return QM31(ac[0] + 2 * bd[0] - bd[1], ac[1] + 2 * bd[1] + bd[0], ad[0] + bc[0], ad[1] + bc[1])

    def inv(self):
        a0, a1, b0, b1 = self.c
        b2 = (b0 * b0 - b1 * b1, 2 * b0 * b1)
        # denom = a^2 - (2 + i) b^2  (CM31)
        d = ((a0 * a0 - a1 * a1) - (2 * b2[0] - b2[1]), 2 * a0 * a1 - (2 * b2[1] + b2[0]))
        n = m_inv((d[0] * d[0] + d[1] * d[1]) % P)
        di = (d[0] * n % P, -d[1] * n % P)
        return QM31(a0 * di[0] - a1 * di[1], a0 * di[1] + a1 * di[0], -(b0 * di[0] - b1 * di[1]), -(b0 * di[1] + b1 * di[0]))

    def __eq__(self, o):
        return self.c == _q(o).c

    def __hash__(self):
        return hash(self.c)

    def is_zero(self):
        return self.c == (0, 0, 0, 0)

    def words(self):
        return list(self.c)


def _q(x):
    return x if isinstance(x, QM31) else QM31(x)


ONE = QM31(1)

# ---- circle group ------------------------------------------------------------------------------------------------------
_GEN = (2, 1268011823)
_MASK = (1 << 31) - 1


def _pt_add(p, q):
    return ((p[0] * q[0] - p[1] * q[1]) % P, (p[0] * q[1] + p[1] * q[0]) % P)


_POW2 = [_GEN]
for _ in range(30):
    _POW2.append(_pt_add(_POW2[-1], _POW2[-1]))


def index_to_point(idx: int):
    """CirclePointIndex::to_point: idx * G."""
    idx &= _MASK
    r = (1, 0)
    j = 0
    while idx:
        if idx & 1:
            r = _pt_add(r, _POW2[j])
        idx >>= 1
        j += 1
    return r


def subgroup_gen(log_size: int) -> int:
    return 1 << (31 - log_size)


def qpt_add(p, q):
    """Point addition with QM31 coordinates."""
    return (p[0] * q[0] - p[1] * q[1], p[0] * q[1] + p[1] * q[0])


def bit_reverse(i: int, bits: int) -> int:
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else i


# ---- Blake2s + channel ----------------------------------------------------------------------------------------------------
_IV = (0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19)
_SIGMA = ((0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), (14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3),
          (11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4), (7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8),
          (9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13), (2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9),
          (12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11), (13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10),
          (6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5), (10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0))
_M32 = 0xFFFFFFFF


def blake2s_compress(h, m, t0=0, t1=0, f0=0, f1=0):
    """RFC 7693 F, the raw compression stwo's legacy mix_u64 / grind use directly."""
    v = list(h) + list(_IV)
    v[12] ^= t0
    v[13] ^= t1
    v[14] ^= f0
    v[15] ^= f1

    def rotr(x, n):
        return ((x >> n) | (x << (32 - n))) & _M32

    def g(a, b, c, d, x, y):
        v[a] = (v[a] + v[b] + x) & _M32
        v[d] = rotr(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & _M32
        v[b] = rotr(v[b] ^ v[c], 12)
        v[a] = (v[a] + v[b] + y) & _M32
        v[d] = rotr(v[d] ^ v[a], 8)
        v[c] = (v[c] + v[d]) & _M32
        v[b] = rotr(v[b] ^ v[c], 7)

    for r in range(10):
        s = _SIGMA[r]
        g(0, 4, 8, 12, m[s[0]], m[s[1]])
        g(1, 5, 9, 13, m[s[2]], m[s[3]])
        g(2, 6, 10, 14, m[s[4]], m[s[5]])
        g(3, 7, 11, 15, m[s[6]], m[s[7]])
        g(0, 5, 10, 15, m[s[8]], m[s[9]])
        g(1, 6, 11, 12, m[s[10]], m[s[11]])
        g(2, 7, 8, 13, m[s[12]], m[s[13]])
        g(3, 4, 9, 14, m[s[14]], m[s[15]])
    return [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]


def blake2s(data: bytes) -> bytes:
    return hashlib.blake2s(data, digest_size=32).digest()


class Blake2sChannel:
    """stwo core/channel/blake2s.rs; variant "legacy" is the one pinned by the reference's committed proof, "v2" the later
    mixing."""

    def __init__(self, variant: str = "legacy"):
        self.variant = variant
        self.digest = bytes(32)
        self.n_sent = 0

    def _update(self, d):
        self.digest, self.n_sent = d, 0

    def mix_root(self, root: bytes):
        self._update(blake2s(self.digest + root))

    def mix_felts(self, felts):
        self._update(blake2s(self.digest + b"".join(struct.pack("<4I", *f.c) for f in felts)))

    def mix_u64(self, v: int):
        lo, hi = v & _M32, (v >> 32) & _M32
        if self.variant == "legacy":
            self._update(struct.pack("<8I", *blake2s_compress(struct.unpack("<8I", self.digest), [lo, hi] + [0] * 14)))
        else:
            self._update(blake2s(self.digest + struct.pack("<2I", lo, hi)))

    def draw_random_bytes(self) -> bytes:
        data = self.digest + self.n_sent.to_bytes(32, "little") + (b"\0" if self.variant == "v2" else b"")
        self.n_sent += 1
        return blake2s(data)

    def draw_base_felts(self):
        while True:
            u = struct.unpack("<8I", self.draw_random_bytes())
            if all(x < 2 * P for x in u):
                return [x % P for x in u]

    def draw_secure_felt(self) -> QM31:
        return QM31(*self.draw_base_felts()[:4])

    def draw_secure_felts(self, n: int):
        out, pool = [], []
        while len(out) < n:
            if len(pool) < 4:
                pool += self.draw_base_felts()
            out.append(QM31(*pool[:4]))
            pool = pool[4:]
        return out
