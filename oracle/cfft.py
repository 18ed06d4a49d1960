"""Circle FFT over M31 (oracle; test infrastructure only).

Restates stwo ``prover/backend/cpu/circle.rs`` @0790eba: ``interpolate``,
``evaluate``, ``eval_at_point``, ``precompute_twiddles`` (reference call sites:
crates/prover/src/prover.rs:38-42,57,59,179,298 and every
``tree_builder.extend_evals`` in crates/air/src/components/*/witness.rs,
e.g. add/witness.rs:51,164).

Layout: evaluations are stored bit-reversed: ``values[j] = f(domain.at(bitrev(j)))``.
Coefficient index bits (LSB first) select ``y, x, pi(x), pi^2(x), ...``.

Butterfly network (n = log size, layers applied to a flat array):
  evaluate   : for i = n-1 .. 1 (line layers, twiddle x), then i = 0 (twiddle y)
               idx0 = (h << (i+1)) + l, idx1 = idx0 + (1 << i), l < 2^i
               (v0, v1) <- (v0 + v1*t, v0 - v1*t)
  interpolate: i = 0 first, then 1 .. n-1, (v0, v1) <- (v0 + v1, (v0 - v1) * t^-1), then * 2^-n
"""
from __future__ import annotations

import numpy as np

from .circle import CircleDomain, Coset, bit_reverse_indices, double_x
from .fields import P, U64, QM31, m_add, m_inv, m_mul, m_sub


def layer_twiddles(domain: CircleDomain):
    """Twiddles for each layer i=0..n-1 of a domain of log size n.

    tw[0][h], h < 2^(n-1): y of half_coset.at(bitrev(h, n-1))      (circle layer)
    tw[i][h], h < 2^(n-1-i): x of (2^(i-1) * half_coset).at(bitrev(h, n-1-i))  (line layers)

    This equals stwo's twiddle tree slices (``slow_precompute_twiddles`` +
    ``domain_line_twiddles_from_tree`` + ``circle_twiddles_from_line_twiddles``).
    """
    n = domain.log_size
    tws = []
    coset = domain.half_coset
    xs, ys = coset.points()
    br = bit_reverse_indices(n - 1)
    tws.append(ys[br] if n > 1 else ys)
    for i in range(1, n):
        m = n - 1 - i  # log2(#twiddles)
        xs, _ = coset.points()
        first_half = xs[: 1 << m]
        tws.append(first_half[bit_reverse_indices(m)])
        coset = coset.double() if coset.log_size > 0 else coset
    return tws


def evaluate(coeffs: np.ndarray, domain: CircleDomain) -> np.ndarray:
    """coeffs[..., 2^k] (k <= n) -> evaluations[..., 2^n] on `domain` (bit-reversed order).

    Zero-extends like stwo ``CirclePoly::extend`` then runs the forward CFFT.
    """
    n = domain.log_size
    N = 1 << n
    c = np.asarray(coeffs, dtype=U64)
    lead = c.shape[:-1]
    v = np.zeros(lead + (N,), dtype=U64)
    v[..., : c.shape[-1]] = c
    tws = layer_twiddles(domain)
    for i in range(n - 1, -1, -1):
        t = tws[i]  # 2^(n-1-i)
        w = v.reshape(lead + (N >> (i + 1), 2, 1 << i))
        v0 = w[..., 0, :]
        v1 = w[..., 1, :]
        tmp = m_mul(v1, t[:, None])
        w0 = m_add(v0, tmp)
        w1 = m_sub(v0, tmp)
        w[..., 0, :] = w0
        w[..., 1, :] = w1
    return v


def interpolate(values: np.ndarray, domain: CircleDomain) -> np.ndarray:
    """evaluations[..., 2^n] (bit-reversed) -> coefficients[..., 2^n]."""
    n = domain.log_size
    N = 1 << n
    v = np.array(values, dtype=U64, copy=True)
    assert v.shape[-1] == N
    lead = v.shape[:-1]
    tws = layer_twiddles(domain)
    for i in range(0, n):
        it = m_inv(tws[i])
        w = v.reshape(lead + (N >> (i + 1), 2, 1 << i))
        v0 = w[..., 0, :].copy()
        v1 = w[..., 1, :].copy()
        w[..., 0, :] = m_add(v0, v1)
        w[..., 1, :] = m_mul(m_sub(v0, v1), it[:, None])
    inv_n = pow(N, P - 2, P)
    return m_mul(v, inv_n)


def eval_at_point(coeffs: np.ndarray, px: QM31, py: QM31) -> QM31:
    """stwo CpuBackend::eval_at_point + utils::fold. coeffs[2^n] M31; point in QM31."""
    c = np.asarray(coeffs, dtype=U64)
    n = int(c.shape[-1]).bit_length() - 1
    if n == 0:
        return QM31(int(c[0]), 0, 0, 0)
    mappings = [py]
    x = px
    for _ in range(1, n):
        mappings.append(x)
        x = double_x(x)
    # fold from the LSB factor (y) upwards: pairs (lo, hi) adjacent.
    cur = QM31.from_m31(c)
    for f in mappings:  # f pairs elements at stride 1 of the current array
        lo = QM31(*[comp[0::2] for comp in cur.c])
        hi = QM31(*[comp[1::2] for comp in cur.c])
        cur = lo + hi * f
    return QM31(*[int(comp[0]) for comp in cur.c])


def eval_at_point_m31(coeffs: np.ndarray, x: int, y: int) -> int:
    """Definition-level evaluation at a base-field point (for O(N^2) checks)."""
    c = [int(t) for t in coeffs]
    n = len(c).bit_length() - 1
    mappings = [y]
    for _ in range(1, n):
        mappings.append(x)
        x = int(double_x(x))
    cur = c
    for f in mappings:
        cur = [(cur[2 * k] + cur[2 * k + 1] * f) % P for k in range(len(cur) // 2)]
    return cur[0]


def precompute_twiddles(root: Coset) -> np.ndarray:
    """stwo ``slow_precompute_twiddles``: per layer, x of the first half of the
    (k-times doubled) coset, bit-reversed; layers concatenated; pad with 1."""
    out = []
    coset = root
    for _ in range(root.log_size):
        xs, _ = coset.points()
        half = xs[: coset.size() // 2]
        out.append(half[bit_reverse_indices(coset.log_size - 1)])
        coset = coset.double()
    out.append(np.array([1], dtype=U64))
    return np.concatenate(out)
