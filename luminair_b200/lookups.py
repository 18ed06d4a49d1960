"""Lookup-table layouts and preprocessed LUT columns of the circuit settings - host-side data format next to the path
(SURVEY 8 f4), not a compute path.

``LookupLayout`` mirrors crates/air/src/preprocessed.rs:41-116 (sorted, disjoint value ranges; ``find_index``);
``lut_columns`` generates the two preprocessed columns of a function table exactly the way the reference does on the host
(preprocessed.rs:351-383 sin, :436-462 exp2, :519-545 log2): ``Fixed::from_f64(f(Fixed(v).to_f64()))`` with Rust's
``f64::sin / exp2 / log2`` - the platform libm, reached here through Python's ``math`` module (also libm), one element at a
time, NOT numpy's SIMD routines, whose last-ulp differences could flip a ``Fixed<12>`` rounding.  ``Fixed::from_f64`` is
numerair's (un-vendored, rev 11d1d26): restated as ``(v * 2^12).round()`` with Rust's ``f64::round`` (ties away from zero)
- parity unpinned.
"""
from __future__ import annotations

import math

import numpy as np

P = (1 << 31) - 1
FP_SCALE = 1 << 12  # numerair Fixed<12>, crates/air/src/lib.rs:23
RANGE_CHECK_BITS = 8  # RangeCheckLookup<1> over the 8-bit limbs of less_than (less_than/component.rs:105-133)


def round_half_away(x) -> np.ndarray:
    """Rust ``f64::round``: to nearest, ties away from zero (numpy's ``round`` ties to even)."""
    x = np.asarray(x, dtype=np.float64)
    return (np.sign(x) * np.floor(np.abs(x) + 0.5)).astype(np.int64)


def to_fixed(x) -> np.ndarray:
    """f32/f64 -> Fixed<12> raw value (``Fixed::from_f64``, crates/graph/src/op/prim.rs:169), as int64."""
    return round_half_away(np.asarray(x, dtype=np.float64) * FP_SCALE)


def _libm(fn):
    def f(v: float) -> float:
        try:
            return fn(v)
        except (ValueError, OverflowError):  # log2 of a non-positive value: Rust gives NaN / -inf
            return float("-inf") if v == 0.0 else float("nan")
    return f


LUT_FUNCS = {"sin": _libm(math.sin), "exp2": _libm(lambda v: math.exp2(v) if hasattr(math, "exp2") else 2.0 ** v),
             "log2": _libm(math.log2)}


def lut_function(name: str, raw_values) -> np.ndarray:
    """f over raw Fixed<12> values -> raw Fixed<12> results (NaN and -inf map to 0, as ``as i64`` saturating casts of the
    reference would never be reached for in-range tables; kept total for padding rows)."""
    fn = LUT_FUNCS[name]
    out = np.empty(len(raw_values), dtype=np.float64)
    for i, v in enumerate(np.asarray(raw_values, dtype=np.int64).tolist()):
        out[i] = fn(v / FP_SCALE)
    out = np.nan_to_num(out, nan=0.0, neginf=0.0)
    return round_half_away(out * FP_SCALE)


class LookupLayout:
    """crates/air/src/preprocessed.rs:41-116: sorted, disjoint value ranges [(lo, hi)] (inclusive raw Fixed<12>
    values); entry i of the table is the i-th value of the concatenated ranges."""

    def __init__(self, ranges):
        self.ranges = sorted((int(lo), int(hi)) for lo, hi in ranges)
        count = sum(hi - lo + 1 for lo, hi in self.ranges)
        # calculate_log_size, crates/air/src/utils.rs:22-27
        self.log_size = max(((count + 15) >> 4) - 1, 0).bit_length() + 4
        self.values = np.concatenate([np.arange(lo, hi + 1, dtype=np.int64) for lo, hi in self.ranges])

    @staticmethod
    def covering(values, pad: int = 0):
        """One range covering the given raw values (the reference derives ranges from a calibration run,
        crates/graph/src/graph.rs gen_circuit_settings)."""
        v = np.asarray(values, dtype=np.int64)
        return LookupLayout([(int(v.min()) - pad, int(v.max()) + pad)])

    def find_index(self, target: np.ndarray) -> np.ndarray:
        idx = np.searchsorted(self.values, target)
        if np.any(idx >= self.values.size) or np.any(self.values[np.minimum(idx, self.values.size - 1)] != target):
            raise ValueError("Value should fit in range.")
        return idx


def lut_columns(name: str, layout: LookupLayout):
    """Preprocessed LUT columns: column 0 = the input values, column 1 = f(value), zero beyond the enumerated values.
    -> [(id, values[2^log_size])]"""
    n = 1 << layout.log_size
    c0 = np.zeros(n, dtype=np.int64)
    c1 = np.zeros(n, dtype=np.int64)
    c0[: layout.values.size] = layout.values
    c1[: layout.values.size] = lut_function(name, layout.values)
    return [(f"{name}_lut_0", (c0 % P).astype(np.uint32)), (f"{name}_lut_1", (c1 % P).astype(np.uint32))]


def range_check_column(n_bits: int = RANGE_CHECK_BITS):
    """preprocessed.rs:233-248,289-307: the enumeration 0 .. 2^n_bits - 1 (one segment)."""
    log_size = max(n_bits, 4)
    col = np.zeros(1 << log_size, dtype=np.uint32)
    col[: 1 << n_bits] = np.arange(1 << n_bits, dtype=np.uint32)
    return [(f"range_check_{n_bits}_column_0", col)]
