"""DEEP / FRI quotients (oracle; test infrastructure only).

Restates stwo ``core/pcs/quotients.rs`` (ColumnSampleBatch, quotient_constants,
accumulate_row_quotients, fri_answers) and ``prover/backend/cpu/quotients.rs``
(accumulate_quotients) @0790eba; reached from the reference through
``prover::prove`` -> ``prove_values`` (crates/prover/src/prover.rs:311-312).
"""
from __future__ import annotations

import numpy as np

from .circle import CanonicCoset
from .fields import CM31, P, U64, QM31, m_neg


def _pt_key(pt):
    return pt[0].tup() + pt[1].tup()


def column_sample_batches(samples):
    """samples[col] = [(point, value)] -> list of (point, [(col_idx, value)]) ordered by
    point (stwo uses a BTreeMap keyed by CirclePoint<SecureField>: x then y, each
    QM31 compared as its 4 u32 coordinates)."""
    groups = {}
    pts = {}
    for ci, col in enumerate(samples):
        for pt, val in col:
            k = _pt_key(pt)
            groups.setdefault(k, []).append((ci, val))
            pts[k] = pt
    return [(pts[k], groups[k]) for k in sorted(groups)]


def quotient_constants(batches, random_coeff: QM31):
    """-> (line_coeffs[batch][j] = (a,b,c), batch_random_coeffs[batch])."""
    line_coeffs = []
    batch_rc = []
    for pt, cols in batches:
        alpha = QM31(1, 0, 0, 0)
        lc = []
        for _, value in cols:
            alpha = alpha * random_coeff
            a = value.complex_conjugate() - value
            c = pt[1].complex_conjugate() - pt[1]
            b = value * c - a * pt[1]
            lc.append((alpha * a, alpha * b, alpha * c))
        line_coeffs.append(lc)
        batch_rc.append(random_coeff.pow(len(cols)))
    return line_coeffs, batch_rc


def denominator_inverses(batches, dx, dy):
    """CM31 1/((Pr.x - x) * Pi.y - (Pr.y - y) * Pi.x) per batch; dx, dy ints or arrays."""
    out = []
    for pt, _ in batches:
        prx, pry = pt[0].lo(), pt[1].lo()
        pix, piy = pt[0].hi(), pt[1].hi()
        d = (prx - dx) * piy - (pry - dy) * pix
        out.append(d.inv())
    return out


def accumulate_row_quotients(batches, consts, values_at_row, dx, dy) -> QM31:
    """values_at_row[col] int or array (all columns of this log size); dx, dy domain point."""
    line_coeffs, batch_rc = consts
    dinv = denominator_inverses(batches, dx, dy)
    acc = None
    for (pt, cols), lcs, brc, di in zip(batches, line_coeffs, batch_rc, dinv):
        num = None
        for (ci, _), (a, b, c) in zip(cols, lcs):
            value = c * values_at_row[ci]
            linear = a * dy + b
            t = value - linear
            num = t if num is None else num + t
        q = num.mul_cm31(di)
        acc = q if acc is None else acc * brc + q
    return acc


def accumulate_quotients(log_size, columns, samples, random_coeff: QM31) -> QM31:
    """Prover side: QM31 column (arrays of 2^log_size, storage order)."""
    dom = CanonicCoset(log_size).circle_domain()
    xs, ys = dom.points_bitrev()
    batches = column_sample_batches(samples)
    consts = quotient_constants(batches, random_coeff)
    cols = [np.asarray(c, dtype=U64) for c in columns]
    return accumulate_row_quotients(batches, consts, cols, xs, ys)


def compute_fri_quotients(columns, samples, random_coeff: QM31):
    """columns: flat list of LDE columns (any sizes), samples: flat list per column.
    -> list of (log_size, QM31 column) by descending log size."""
    order = sorted(range(len(columns)), key=lambda i: -len(columns[i]))
    out = []
    i = 0
    while i < len(order):
        n = len(columns[order[i]])
        grp = []
        while i < len(order) and len(columns[order[i]]) == n:
            grp.append(order[i])
            i += 1
        log_size = n.bit_length() - 1
        out.append((log_size, accumulate_quotients(log_size, [columns[g] for g in grp], [samples[g] for g in grp], random_coeff)))
    return out
