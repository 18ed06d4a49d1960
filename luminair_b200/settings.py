"""``CircuitSettings`` (crates/air/src/settings.rs:21-122): the lookup-table configuration a proof is made and verified
against - per LUT its value ranges (``LookupLayout``, crates/air/src/preprocessed.rs:32-57) and the multiplicity column
filled by ``gen_trace`` (``AtomicMultiplicityColumn``, crates/air/src/utils.rs:66-92) - with the reference's bincode 1.3
wire format (derive(Serialize) field order; little-endian; ``Option`` = one tag byte; ``Vec`` = u64 length):

    CircuitSettings { lookups: Lookups { sin, exp2, log2: Option<{ layout: { ranges: Vec<(i64, i64)>, log_size: u32 },
                                                                multiplicities: { data: Vec<u32> } }>,
                                         range_check: Option<{ layout: { ranges: [u32; 1], log_size: u32 },
                                                               multiplicities: { data: Vec<u32> } }> } }

The reference commits one settings file (ui/demo/public/settings, a single 0x00 byte: the one-``Option`` schema of the
revision that produced the demo proof); the 17-component schema above is restated from the in-tree struct definitions
(components/lookups/mod.rs:19-29, lookups/exp2/mod.rs:23-27, lookups/range_check/mod.rs:23-35) - parity unpinned.

Host-side data format next to the path: ``prove(pie, settings=...)`` derives the preprocessed LUT columns from it exactly as
prover.rs:52-55 does (``lookups_to_preprocessed_column``).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .lookups import RANGE_CHECK_BITS, LookupLayout, lut_columns, range_check_column

LUT_ORDER = ("sin", "exp2", "log2")  # field order of `Lookups`, then range_check


@dataclass
class Lookup:
    """SinLookup / Exp2Lookup / Log2Lookup: value ranges (inclusive raw Fixed<12>) + multiplicities."""
    ranges: List[Tuple[int, int]]
    log_size: int
    multiplicities: np.ndarray  # u32[2^log_size]

    def layout(self) -> LookupLayout:
        lay = LookupLayout(self.ranges)
        if lay.log_size != self.log_size:
            raise ValueError("LookupLayout.log_size does not match its ranges (calculate_log_size)")
        return lay


@dataclass
class RangeCheckLookup:
    """RangeCheckLookup<1>: one bit-length segment."""
    n_bits: int
    log_size: int
    multiplicities: np.ndarray


@dataclass
class CircuitSettings:
    sin: Optional[Lookup] = None
    exp2: Optional[Lookup] = None
    log2: Optional[Lookup] = None
    range_check: Optional[RangeCheckLookup] = None

    # ---- what prove() needs -----------------------------------------------------------------------
    def preprocessed_columns(self):
        """lookups_to_preprocessed_column (preprocessed.rs:181-206): [(id, values)] in sin, exp2, log2, range_check order."""
        cols = []
        for name in LUT_ORDER:
            lk = getattr(self, name)
            if lk is not None:
                cols += lut_columns(name, lk.layout())
        if self.range_check is not None:
            cols += range_check_column(self.range_check.n_bits)
        return cols

    def preprocessed_meta(self):
        return [(cid, len(v).bit_length() - 1) for cid, v in self.preprocessed_columns()]

    @staticmethod
    def from_graph_trace(pie, preprocessed) -> "CircuitSettings":
        """Settings of a ``GraphTrace.finish()`` result: one covering range per LUT, multiplicities from the lookup tables."""
        tables = dict(pie)
        pre = dict(preprocessed)
        s = CircuitSettings()
        for name in LUT_ORDER:
            if f"{name}_lut_0" in pre:
                mult = np.asarray(tables[f"{name}_lookup"], dtype=np.uint32).reshape(-1)
                vals = np.asarray(pre[f"{name}_lut_0"], dtype=np.int64)
                n_vals = int(np.count_nonzero(np.asarray(pre[f"{name}_lut_1"])) or 1)
                # the table enumerates one contiguous range; recover it from the first entry and the multiplicity support
                lo = int(vals[0]) if vals[0] < (1 << 30) else int(vals[0]) - ((1 << 31) - 1)
                count = _covering_count(vals)
                setattr(s, name, Lookup([(lo, lo + count - 1)], mult.size.bit_length() - 1, mult))
        if any(cid.startswith("range_check_") for cid in pre):
            mult = np.asarray(tables["range_check_lookup"], dtype=np.uint32).reshape(-1)
            s.range_check = RangeCheckLookup(RANGE_CHECK_BITS, mult.size.bit_length() - 1, mult)
        return s

    # ---- bincode 1.3 -------------------------------------------------------------------------------
    def to_bincode(self) -> bytes:
        out = bytearray()
        for name in LUT_ORDER:
            lk = getattr(self, name)
            if lk is None:
                out += b"\x00"
                continue
            out += b"\x01" + struct.pack("<Q", len(lk.ranges))
            for lo, hi in lk.ranges:
                out += struct.pack("<qq", lo, hi)
            out += struct.pack("<I", lk.log_size)
            m = np.ascontiguousarray(lk.multiplicities, dtype="<u4")
            out += struct.pack("<Q", m.size) + m.tobytes()
        rc = self.range_check
        if rc is None:
            out += b"\x00"
        else:
            m = np.ascontiguousarray(rc.multiplicities, dtype="<u4")
            out += b"\x01" + struct.pack("<II", rc.n_bits, rc.log_size) + struct.pack("<Q", m.size) + m.tobytes()
        return bytes(out)

    @staticmethod
    def from_bincode(data: bytes) -> "CircuitSettings":
        pos = 0

        def take(fmt):
            nonlocal pos
            vals = struct.unpack_from(fmt, data, pos)
            pos += struct.calcsize(fmt)
            return vals

        def take_u32s():
            nonlocal pos
            (n,) = take("<Q")
            arr = np.frombuffer(data, dtype="<u4", count=n, offset=pos).copy()
            pos += 4 * n
            return arr

        s = CircuitSettings()
        for name in LUT_ORDER:
            (tag,) = take("<B")
            if tag == 0:
                continue
            if tag != 1:
                raise ValueError("bad Option tag")
            (n,) = take("<Q")
            ranges = [take("<qq") for _ in range(n)]
            (log_size,) = take("<I")
            setattr(s, name, Lookup([(int(a), int(b)) for a, b in ranges], int(log_size), take_u32s()))
        (tag,) = take("<B")
        if tag == 1:
            n_bits, log_size = take("<II")
            s.range_check = RangeCheckLookup(int(n_bits), int(log_size), take_u32s())
        elif tag != 0:
            raise ValueError("bad Option tag")
        if pos != len(data):
            raise ValueError("trailing bytes after CircuitSettings")
        return s


    # ---- JSON (serde_json of the same structs, settings.rs:64-122) -----------------------------------
    def to_json(self) -> str:
        """``CircuitSettings::to_json``: {"lookups": {"sin" | "exp2" | "log2": null | {"layout": {"ranges": [[lo, hi], ..],
        "log_size": k}, "multiplicities": {"data": [..]}}, "range_check": null | {"layout": {"ranges": [n_bits], "log_size": k},
        "multiplicities": {"data": [..]}}}}.  ``Range(Fixed, Fixed)`` is a tuple struct of two newtypes over i64
        (preprocessed.rs:35) - a pair of numbers; ``AtomicU32`` serialises as its value.  Parity unpinned (no JSON settings
        file in the reference)."""
        import json

        def lut(lk):
            if lk is None:
                return None
            return {"layout": {"ranges": [[int(lo), int(hi)] for lo, hi in lk.ranges], "log_size": int(lk.log_size)},
                    "multiplicities": {"data": [int(x) for x in np.asarray(lk.multiplicities).reshape(-1)]}}

        rc = self.range_check
        doc = {"lookups": {"sin": lut(self.sin), "exp2": lut(self.exp2), "log2": lut(self.log2),
                           "range_check": None if rc is None else {
                               "layout": {"ranges": [int(rc.n_bits)], "log_size": int(rc.log_size)},
                               "multiplicities": {"data": [int(x) for x in np.asarray(rc.multiplicities).reshape(-1)]}}}}
        return json.dumps(doc, indent=2)

    @staticmethod
    def from_json(text: str) -> "CircuitSettings":
        import json
        lk = json.loads(text)["lookups"]
        s = CircuitSettings()
        for name in LUT_ORDER:
            j = lk.get(name)
            if j is not None:
                setattr(s, name, Lookup([(int(a), int(b)) for a, b in j["layout"]["ranges"]], int(j["layout"]["log_size"]),
                                        np.asarray(j["multiplicities"]["data"], dtype=np.uint32)))
        j = lk.get("range_check")
        if j is not None:
            s.range_check = RangeCheckLookup(int(j["layout"]["ranges"][0]), int(j["layout"]["log_size"]),
                                             np.asarray(j["multiplicities"]["data"], dtype=np.uint32))
        return s


def _covering_count(lut0_values: np.ndarray) -> int:
    """Number of enumerated values of a one-range table: entries are consecutive raw values, the padding is zeros."""
    P = (1 << 31) - 1
    v = np.asarray(lut0_values, dtype=np.int64)
    signed = np.where(v >= (1 << 30), v - P, v)
    steps = np.diff(signed)
    breaks = np.nonzero(steps != 1)[0]
    return int(breaks[0]) + 1 if breaks.size else int(v.size)
