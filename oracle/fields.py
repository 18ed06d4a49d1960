"""M31 / CM31 / QM31 arithmetic (oracle; test infrastructure only).

Restates stwo ``core/fields/{m31,cm31,qm31}.rs`` @0790eba (un-vendored; call
sites in the reference: crates/air/src/components/mod.rs:203-211 (SecureField
claimed sums), crates/air/src/components/mul/component.rs:65 (M31 constants)).

Every function works on Python ints *and* on numpy ``uint64`` arrays, so the
same code serves scalar transcript work and vectorised column work.

  M31  : p = 2^31 - 1
  CM31 : M31[i]/(i^2 + 1)
  QM31 : CM31[u]/(u^2 - (2 + i)); serialised as 4 LE u32 (a.re, a.im, b.re, b.im)
"""
from __future__ import annotations

import numpy as np

P = (1 << 31) - 1
U64 = np.uint64


def _is_arr(x):
    return isinstance(x, np.ndarray)


def m_add(a, b):
    return (a + b) % P


def m_sub(a, b):
    return (a + P - b) % P


def m_neg(a):
    return (P - a) % P


def m_mul(a, b):
    return (a * b) % P


def m_pow(a, e: int):
    """a^e for int or uint64 array a."""
    if _is_arr(a):
        res = np.ones_like(a)
    else:
        res = 1
    base = a
    while e:
        if e & 1:
            res = m_mul(res, base)
        base = m_mul(base, base)
        e >>= 1
    return res


def m_inv(a):
    return m_pow(a, P - 2)


def m_batch_inv(a: np.ndarray) -> np.ndarray:
    """Element-wise inverse (Fermat; the oracle favours clarity over speed)."""
    return m_inv(a)


def m_reduce_u64(x):
    """stwo M31::reduce: ((((x >> 31) + x + 1) >> 31) + x) & p for x < p^2."""
    return x % P


class CM31:
    """a + b*i, components int or uint64 ndarray."""

    __slots__ = ("a", "b")

    def __init__(self, a, b=0):
        self.a = a
        self.b = b

    def __add__(self, o):
        if isinstance(o, CM31):
            return CM31(m_add(self.a, o.a), m_add(self.b, o.b))
        return CM31(m_add(self.a, o), self.b)

    def __sub__(self, o):
        if isinstance(o, CM31):
            return CM31(m_sub(self.a, o.a), m_sub(self.b, o.b))
        return CM31(m_sub(self.a, o), self.b)

    def __neg__(self):
        return CM31(m_neg(self.a), m_neg(self.b))

    def __mul__(self, o):
        if isinstance(o, CM31):
            return CM31(
                m_sub(m_mul(self.a, o.a), m_mul(self.b, o.b)),
                m_add(m_mul(self.a, o.b), m_mul(self.b, o.a)),
            )
        return CM31(m_mul(self.a, o), m_mul(self.b, o))

    def inv(self):
        n = m_inv(m_add(m_mul(self.a, self.a), m_mul(self.b, self.b)))
        return CM31(m_mul(self.a, n), m_mul(m_neg(self.b), n))

    def __eq__(self, o):
        return bool(np.all(self.a == o.a) and np.all(self.b == o.b))


class QM31:
    """(a0 + a1 i) + (b0 + b1 i) u ; components int or uint64 ndarray."""

    __slots__ = ("c",)

    def __init__(self, c0=0, c1=0, c2=0, c3=0):
        self.c = (c0, c1, c2, c3)

    # --- constructors -----------------------------------------------------
    @staticmethod
    def from_m31(x):
        z = x * 0
        return QM31(x, z, z, z)

    @staticmethod
    def from_array(arr: np.ndarray) -> "QM31":
        """arr[..., 4] uint32/uint64 -> QM31 of arrays."""
        arr = arr.astype(U64)
        return QM31(arr[..., 0], arr[..., 1], arr[..., 2], arr[..., 3])

    def to_array(self) -> np.ndarray:
        return np.stack([np.asarray(x, dtype=U64) for x in np.broadcast_arrays(*self.c)], axis=-1).astype(np.uint32)

    @staticmethod
    def from_partial_evals(e):
        """stwo SecureField::from_partial_evals: e0 + e1*i + e2*u + e3*iu."""
        return e[0] + e[1] * QM31(0, 1, 0, 0) + e[2] * QM31(0, 0, 1, 0) + e[3] * QM31(0, 0, 0, 1)

    # --- helpers ----------------------------------------------------------
    def lo(self) -> CM31:
        return CM31(self.c[0], self.c[1])

    def hi(self) -> CM31:
        return CM31(self.c[2], self.c[3])

    @staticmethod
    def from_cm31(a: CM31, b: CM31) -> "QM31":
        return QM31(a.a, a.b, b.a, b.b)

    def tup(self):
        return tuple(int(x) for x in self.c)

    def __getitem__(self, idx) -> "QM31":
        return QM31(*[x[idx] for x in self.c])

    # --- arithmetic -------------------------------------------------------
    def __add__(self, o):
        if isinstance(o, QM31):
            return QM31(*[m_add(x, y) for x, y in zip(self.c, o.c)])
        return QM31(m_add(self.c[0], o), self.c[1], self.c[2], self.c[3])

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, QM31):
            return QM31(*[m_sub(x, y) for x, y in zip(self.c, o.c)])
        return QM31(m_sub(self.c[0], o), self.c[1], self.c[2], self.c[3])

    def __rsub__(self, o):
        return (-self) + o

    def __neg__(self):
        return QM31(*[m_neg(x) for x in self.c])

    def __mul__(self, o):
        if isinstance(o, QM31):
            a, b, c, d = self.lo(), self.hi(), o.lo(), o.hi()
            bd = b * d
            # (2 + i) * bd
            r_bd = CM31(m_sub(m_add(bd.a, bd.a), bd.b), m_add(m_add(bd.b, bd.b), bd.a))
            return QM31.from_cm31(a * c + r_bd, a * d + b * c)
        if isinstance(o, CM31):
            return QM31.from_cm31(self.lo() * o, self.hi() * o)
        return QM31(*[m_mul(x, o) for x in self.c])

    __rmul__ = __mul__

    def mul_cm31(self, o: CM31) -> "QM31":
        return QM31.from_cm31(self.lo() * o, self.hi() * o)

    def square(self):
        return self * self

    def inv(self):
        a, b = self.lo(), self.hi()
        b2 = b * b
        ib2 = CM31(m_neg(b2.b), b2.a)
        denom = a * a - (b2 + b2 + ib2)
        di = denom.inv()
        return QM31.from_cm31(a * di, (-b) * di)

    def pow(self, e: int):
        res = QM31(1, 0, 0, 0)
        base = self
        while e:
            if e & 1:
                res = res * base
            base = base * base
            e >>= 1
        return res

    def complex_conjugate(self):
        """stwo QM31::complex_conjugate: a - b u."""
        return QM31(self.c[0], self.c[1], m_neg(self.c[2]), m_neg(self.c[3]))

    def __eq__(self, o):
        if not isinstance(o, QM31):
            o = QM31(o, 0, 0, 0)
        return all(bool(np.all(np.asarray(x) == np.asarray(y))) for x, y in zip(self.c, o.c))

    def __hash__(self):
        return hash(self.tup())

    def is_zero(self):
        return self == QM31()

    def __repr__(self):
        return f"QM31{self.c}"


ONE = QM31(1, 0, 0, 0)
ZERO = QM31(0, 0, 0, 0)


def qm31_zeros(n: int) -> QM31:
    z = np.zeros(n, dtype=U64)
    return QM31(z.copy(), z.copy(), z.copy(), z.copy())


def as_u64(x) -> np.ndarray:
    return np.asarray(x, dtype=U64)
