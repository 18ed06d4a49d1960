"""Circle group over M31, cosets, domains (oracle; test infrastructure only).

Restates stwo ``core/circle.rs`` and ``core/poly/circle/{canonic,domain}.rs``
@0790eba.  Reference call site: crates/prover/src/prover.rs:38-42
(``CanonicCoset::new(..).circle_domain().half_coset``).

Points of the order-2^31 circle group are handled as *indices* (multiples of
the generator G = (2, 1268011823)), exactly like stwo's ``CirclePointIndex``.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from .fields import P, U64, QM31, m_add, m_mul, m_sub, m_neg

M31_CIRCLE_LOG_ORDER = 31
GEN = (2, 1268011823)
ORDER = 1 << M31_CIRCLE_LOG_ORDER


def pt_add(p, q):
    """(x0 x1 - y0 y1, x0 y1 + y0 x1); works for ints, arrays, QM31."""
    if isinstance(p[0], QM31) or isinstance(q[0], QM31):
        return (p[0] * q[0] - p[1] * q[1], p[0] * q[1] + p[1] * q[0])
    return (m_sub(m_mul(p[0], q[0]), m_mul(p[1], q[1])), m_add(m_mul(p[0], q[1]), m_mul(p[1], q[0])))


def pt_double(p):
    return pt_add(p, p)


def double_x(x):
    """pi(x) = 2x^2 - 1."""
    if isinstance(x, QM31):
        return x * x * 2 - 1
    return m_sub(m_mul(2, m_mul(x, x)), 1)


@lru_cache(maxsize=None)
def _gen_pow2(j: int):
    """G * 2^j."""
    p = GEN
    for _ in range(j):
        p = pt_double(p)
    return (int(p[0]), int(p[1]))


def index_to_point(idx: int):
    """CirclePointIndex::to_point for a scalar index."""
    idx %= ORDER
    res = (1, 0)
    j = 0
    while idx:
        if idx & 1:
            res = pt_add(res, _gen_pow2(j))
        idx >>= 1
        j += 1
    return (int(res[0]), int(res[1]))


def subgroup_gen(log_size: int) -> int:
    return 1 << (M31_CIRCLE_LOG_ORDER - log_size)


class Coset:
    """stwo Coset {initial_index, step_size, log_size}."""

    def __init__(self, initial_index: int, log_size: int):
        self.initial_index = initial_index % ORDER
        self.log_size = log_size
        self.step_size = subgroup_gen(log_size)

    @staticmethod
    def odds(log_size):
        return Coset(subgroup_gen(log_size + 1), log_size)

    @staticmethod
    def half_odds(log_size):
        return Coset(subgroup_gen(log_size + 2), log_size)

    @staticmethod
    def subgroup(log_size):
        return Coset(0, log_size)

    def size(self):
        return 1 << self.log_size

    def double(self):
        assert self.log_size > 0
        return Coset(self.initial_index * 2, self.log_size - 1)

    def index_at(self, i: int) -> int:
        return (self.initial_index + self.step_size * i) % ORDER

    def at(self, i: int):
        return index_to_point(self.index_at(i))

    def points(self):
        """All points in natural order as (x[], y[]) uint64 arrays (built by doubling)."""
        n = self.size()
        xs = np.empty(n, dtype=U64)
        ys = np.empty(n, dtype=U64)
        x0, y0 = index_to_point(self.initial_index)
        xs[0], ys[0] = x0, y0
        for j in range(self.log_size):
            sx, sy = index_to_point(self.step_size << j)
            h = 1 << j
            nx, ny = pt_add((xs[:h], ys[:h]), (sx, sy))
            xs[h : 2 * h], ys[h : 2 * h] = nx, ny
        return xs, ys

    def step_point(self):
        return index_to_point(self.step_size)

    def __eq__(self, o):
        return (self.initial_index, self.log_size) == (o.initial_index, o.log_size)


class CircleDomain:
    """half_coset + its conjugate; ``at(i)`` as in stwo CircleDomain::at."""

    def __init__(self, half_coset: Coset):
        self.half_coset = half_coset

    @property
    def log_size(self):
        return self.half_coset.log_size + 1

    def size(self):
        return 1 << self.log_size

    def index_at(self, i: int) -> int:
        h = self.half_coset.size()
        if i < h:
            return self.half_coset.index_at(i)
        return (-self.half_coset.index_at(i - h)) % ORDER

    def at(self, i: int):
        return index_to_point(self.index_at(i))

    def points(self):
        hx, hy = self.half_coset.points()
        return np.concatenate([hx, hx]), np.concatenate([hy, m_neg(hy)])

    def points_bitrev(self):
        """Points in storage (bit-reversed) order: entry j is at(bitrev(j))."""
        xs, ys = self.points()
        br = bit_reverse_indices(self.log_size)
        return xs[br], ys[br]


class CanonicCoset:
    def __init__(self, log_size: int):
        assert log_size > 0
        self.log_size = log_size
        self.coset = Coset.odds(log_size)

    def circle_domain(self) -> CircleDomain:
        return CircleDomain(Coset.half_odds(self.log_size - 1))

    def half_coset(self) -> Coset:
        return Coset.half_odds(self.log_size - 1)

    def step_size(self) -> int:
        return self.coset.step_size

    def step(self):
        return index_to_point(self.coset.step_size)


class LineDomain:
    def __init__(self, coset: Coset):
        self.coset = coset

    @property
    def log_size(self):
        return self.coset.log_size

    def size(self):
        return self.coset.size()

    def at(self, i: int) -> int:
        return self.coset.at(i)[0]

    def double(self):
        return LineDomain(self.coset.double())

    def xs_bitrev(self):
        xs, _ = self.coset.points()
        return xs[bit_reverse_indices(self.log_size)]


@lru_cache(maxsize=32)
def bit_reverse_indices(log_size: int) -> np.ndarray:
    n = 1 << log_size
    idx = np.arange(n, dtype=np.int64)
    res = np.zeros(n, dtype=np.int64)
    for b in range(log_size):
        res |= ((idx >> b) & 1) << (log_size - 1 - b)
    return res


def bit_reverse_index(i: int, log_size: int) -> int:
    if log_size == 0:
        return i
    return int(format(i, f"0{log_size}b")[::-1], 2)


def coset_vanishing(coset: Coset, p):
    """stwo core/constraints.rs::coset_vanishing: x-coordinate after log_size-1 doublings
    of (p - initial + step/2)."""
    half_step = index_to_point(coset.step_size >> 1)
    init_neg = index_to_point(-coset.initial_index)
    q = pt_add(pt_add(p, _lift(init_neg, p)), _lift(half_step, p))
    x = q[0]
    for _ in range(1, coset.log_size):
        x = double_x(x)
    return x


def _lift(pt, like):
    if isinstance(like[0], QM31):
        return (QM31(pt[0], 0, 0, 0), QM31(pt[1], 0, 0, 0))
    return pt
