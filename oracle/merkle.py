"""Blake2s Merkle commitment: mixed-height commit / decommit / verify
(oracle; test infrastructure only).

Restates stwo ``core/vcs/blake2_merkle.rs`` (hash_node), ``prover/vcs/prover.rs``
(MerkleProver::commit / decommit) and ``core/vcs/verifier.rs`` @0790eba.
Reference call sites: every ``tree_builder.commit(channel)``,
crates/prover/src/prover.rs:59,179,298.

hash_node(children, column_values) = Blake2s-256(left || right || values as LE u32)
[artifact-verified].  A tree over columns of several sizes has one layer per
log size from the largest down to 0; a column of log size s is absorbed by the
nodes of layer s.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import blake2s


def hash_node(children, values) -> bytes:
    data = b""
    if children is not None:
        data += children[0] + children[1]
    data += np.asarray(values, dtype="<u4").tobytes()
    return blake2s.hash(data)


def _sorted_columns(columns):
    """Stable sort by length descending (itertools ``sorted_by_key(Reverse(len))``)."""
    return sorted(columns, key=lambda c: -len(c))


class MerkleProver:
    def __init__(self, layers):
        self.layers = layers  # layers[k] = list of 2^k hashes (bytes)

    @staticmethod
    def commit(columns) -> "MerkleProver":
        columns = [np.asarray(c, dtype=np.uint32) for c in columns]
        if not columns:
            return MerkleProver([[blake2s.hash(b"")]])
        cols = _sorted_columns(columns)
        max_log = len(cols[0]).bit_length() - 1
        layers = []
        prev = None
        for log in range(max_log, -1, -1):
            layer_cols = [c for c in cols if len(c) == (1 << log)]
            mat = np.stack(layer_cols, axis=1).astype("<u4") if layer_cols else None
            cur = []
            for i in range(1 << log):
                data = b""
                if prev is not None:
                    data += prev[2 * i] + prev[2 * i + 1]
                if mat is not None:
                    data += mat[i].tobytes()
                cur.append(blake2s.hash(data))
            layers.append(cur)
            prev = cur
        layers.reverse()
        return MerkleProver(layers)

    def root(self) -> bytes:
        return self.layers[0][0]

    def decommit(self, queries_per_log_size, columns):
        """-> (queried_values flat list, hash_witness list, column_witness list)."""
        columns = [np.asarray(c, dtype=np.uint32) for c in columns]
        cols = _sorted_columns(columns)
        queried_values, hash_witness, column_witness = [], [], []
        last_layer_queries = []
        for log in range(len(self.layers) - 1, -1, -1):
            layer_cols = [c for c in cols if len(c) == (1 << log)]
            prev_hashes = self.layers[log + 1] if log + 1 < len(self.layers) else None
            prev_q = list(last_layer_queries)
            col_q = list(queries_per_log_size.get(log, []))
            pi = ci = 0
            total = []
            while pi < len(prev_q) or ci < len(col_q):
                cands = []
                if pi < len(prev_q):
                    cands.append(prev_q[pi] // 2)
                if ci < len(col_q):
                    cands.append(col_q[ci])
                node = min(cands)
                if prev_hashes is not None:
                    if pi < len(prev_q) and prev_q[pi] == 2 * node:
                        pi += 1
                    else:
                        hash_witness.append(prev_hashes[2 * node])
                    if pi < len(prev_q) and prev_q[pi] == 2 * node + 1:
                        pi += 1
                    else:
                        hash_witness.append(prev_hashes[2 * node + 1])
                vals = [int(c[node]) for c in layer_cols]
                if ci < len(col_q) and col_q[ci] == node:
                    ci += 1
                    queried_values.extend(vals)
                else:
                    column_witness.extend(vals)
                total.append(node)
            last_layer_queries = total
        return queried_values, hash_witness, column_witness


class MerkleVerificationError(Exception):
    pass


def verify(root: bytes, column_log_sizes, queries_per_log_size, queried_values, hash_witness, column_witness):
    """core/vcs/verifier.rs MerkleVerifier::verify. Raises on failure."""
    if not column_log_sizes:
        return
    max_log = max(column_log_sizes)
    n_cols = {}
    for s in column_log_sizes:
        n_cols[s] = n_cols.get(s, 0) + 1
    qv = iter(queried_values)
    hw = iter(hash_witness)
    cw = iter(column_witness)
    last = None  # list of (index, hash)
    for log in range(max_log, -1, -1):
        n_in_layer = n_cols.get(log, 0)
        prev = list(last) if last is not None else []
        col_q = list(queries_per_log_size.get(log, []))
        pi = ci = 0
        total = []
        while pi < len(prev) or ci < len(col_q):
            cands = []
            if pi < len(prev):
                cands.append(prev[pi][0] // 2)
            if ci < len(col_q):
                cands.append(col_q[ci])
            node = min(cands)
            children = None
            if last is not None:
                try:
                    if pi < len(prev) and prev[pi][0] == 2 * node:
                        left = prev[pi][1]
                        pi += 1
                    else:
                        left = next(hw)
                    if pi < len(prev) and prev[pi][0] == 2 * node + 1:
                        right = prev[pi][1]
                        pi += 1
                    else:
                        right = next(hw)
                except StopIteration:
                    raise MerkleVerificationError("WitnessTooShort")
                children = (left, right)
            if ci < len(col_q) and col_q[ci] == node:
                ci += 1
                src, err = qv, "TooFewQueriedValues"
            else:
                src, err = cw, "WitnessTooShort"
            vals = []
            for _ in range(n_in_layer):
                try:
                    vals.append(next(src))
                except StopIteration:
                    raise MerkleVerificationError(err)
            total.append((node, hash_node(children, vals)))
        last = total
    for it, name in ((hw, "WitnessTooLong"), (qv, "TooManyQueriedValues"), (cw, "WitnessTooLong")):
        try:
            next(it)
            raise MerkleVerificationError(name)
        except StopIteration:
            pass
    if len(last) != 1 or last[0][1] != root:
        raise MerkleVerificationError("RootMismatch")
