"""FRI over the circle / line domains (oracle; test infrastructure only).

Restates stwo ``core/fri.rs`` (FriVerifier, Queries handling, fold formulas) and
``prover/fri.rs`` + ``prover/backend/cpu/fri.rs`` (FriProver::commit/decommit,
fold_line, fold_circle_into_line) @0790eba.  Reached from the reference through
``prover::prove`` (crates/prover/src/prover.rs:311-312) and
``stwo::core::verifier::verify`` (crates/verifiers/rust/src/verifier.rs:126-132).

FOLD_STEP = 1, CIRCLE_TO_LINE_FOLD_STEP = 1.
"""
from __future__ import annotations

import struct

import numpy as np

from . import merkle
from .circle import CanonicCoset, CircleDomain, Coset, LineDomain, bit_reverse_index, bit_reverse_indices, double_x
from .fields import P, U64, QM31, m_inv, m_mul
from .proof import FriLayerProof, FriProof, MerkleDecommitment


class FriVerificationError(Exception):
    pass


# ---- queries ---------------------------------------------------------------
def generate_queries(channel, log_domain_size: int, n_queries: int):
    qs = set()
    cnt = 0
    mask = (1 << log_domain_size) - 1
    while True:
        rb = channel.draw_random_bytes()
        for (w,) in struct.iter_unpack("<I", rb):
            qs.add(w & mask)
            cnt += 1
            if cnt == n_queries:
                return sorted(qs)


def fold_queries(positions, n_folds: int):
    out = []
    for q in positions:
        f = q >> n_folds
        if not out or out[-1] != f:
            out.append(f)
    return out


def query_positions_by_log_size(queries, log_domain_size, column_log_sizes):
    return {s: fold_queries(queries, log_domain_size - s) for s in sorted(set(column_log_sizes))}


# ---- folds (vectorised over a QM31 column) ----------------------------------
def fold_line(values: QM31, domain: LineDomain, alpha: QM31) -> QM31:
    """values: QM31 arrays of domain.size() in bit-reversed order -> half size."""
    n = domain.size()
    xs = domain.xs_bitrev()[0::2]  # x at storage index 2i
    xinv = m_inv(xs)
    f_x = values[0::2]
    f_nx = values[1::2]
    f0 = f_x + f_nx
    f1 = (f_x - f_nx) * xinv
    return f0 + alpha * f1


def fold_circle_into_line(dst: QM31, src: QM31, src_domain: CircleDomain, alpha: QM31) -> QM31:
    _, ys = src_domain.points_bitrev()
    yinv = m_inv(ys[0::2])
    f_p = src[0::2]
    f_np = src[1::2]
    f0 = f_p + f_np
    f1 = (f_p - f_np) * yinv
    f_prime = alpha * f1 + f0
    return dst * (alpha * alpha) + f_prime


def line_interpolate(values: QM31, domain: LineDomain):
    """LineEvaluation::interpolate -> ordered coefficients (list of QM31 scalars)."""
    n = domain.size()
    log_n = domain.log_size
    br = bit_reverse_indices(log_n)
    vals = [values[int(br[i])] for i in range(n)]  # natural order
    vals = [QM31(*[int(c) for c in v.c]) for v in vals]
    dom = domain
    size = n
    while size > 1:
        xs = [dom.at(i) for i in range(size // 2)]
        for start in range(0, n, size):
            for i in range(size // 2):
                l, r = vals[start + i], vals[start + size // 2 + i]
                xinv = pow(xs[i], P - 2, P)
                vals[start + i] = l + r
                vals[start + size // 2 + i] = (l - r) * xinv
        dom = dom.double()
        size //= 2
    inv_n = pow(n, P - 2, P)
    vals = [v * inv_n for v in vals]
    # into_ordered_coefficients: bit reverse
    return [vals[int(br[i])] for i in range(n)]


def line_poly_eval(coeffs_ordered, x: QM31) -> QM31:
    """LinePoly::eval_at_point on ordered coefficients (c0 + c1 x + c2 pi(x) + c3 x pi(x) ...)."""
    n = len(coeffs_ordered)
    log_n = n.bit_length() - 1
    doublings = []
    for _ in range(log_n):
        doublings.append(x)
        x = double_x(x)
    # ordered coefficient k = prod over bits b of k of doublings[b]
    acc = QM31()
    for k, c in enumerate(coeffs_ordered):
        term = c
        for b in range(log_n):
            if (k >> b) & 1:
                term = term * doublings[b]
        acc = acc + term
    return acc


# ---- prover ------------------------------------------------------------------
def _coord_columns(q: QM31):
    return [np.asarray(c, dtype=np.uint32) for c in q.c]


def _decommit_positions_and_witness(column: QM31, query_positions, fold_step):
    positions, witness = [], []
    i = 0
    qp = list(query_positions)
    while i < len(qp):
        j = i
        while j < len(qp) and (qp[j] >> fold_step) == (qp[i] >> fold_step):
            j += 1
        subset = qp[i:j]
        start = (subset[0] >> fold_step) << fold_step
        k = 0
        for pos in range(start, start + (1 << fold_step)):
            positions.append(pos)
            if k < len(subset) and subset[k] == pos:
                k += 1
                continue
            witness.append(QM31(*[int(c[pos]) for c in column.c]))
        i = j
    return positions, witness


class FriProver:
    """columns: list of (log_size, QM31 column) with strictly decreasing sizes."""

    def __init__(self, channel, config, columns):
        self.config = config
        self.columns = columns
        assert all(a[0] > b[0] for a, b in zip(columns, columns[1:]))
        # first layer: one mixed-height tree over all coordinate columns
        coord = []
        for _, q in columns:
            coord.extend(_coord_columns(q))
        self.first_tree = merkle.MerkleProver.commit(coord)
        channel.mix_root(self.first_tree.root())
        # inner layers
        folding_alpha = channel.draw_secure_felt()
        self.circle_alpha = folding_alpha
        first_log = columns[0][0] - 1
        dom = LineDomain(Coset.half_odds(first_log))
        n = dom.size()
        z = np.zeros(n, dtype=U64)
        layer_eval = QM31(z, z.copy(), z.copy(), z.copy())
        self.inner = []  # (domain, QM31 eval, MerkleProver)
        self.alphas = []
        ci = 0
        last_size = 1 << (config.log_last_layer_degree_bound + config.log_blowup_factor)
        while dom.size() > last_size:
            while ci < len(columns) and (1 << (columns[ci][0] - 1)) == dom.size():
                log_c, col = columns[ci]
                layer_eval = fold_circle_into_line(layer_eval, col, CanonicCoset(log_c).circle_domain(), folding_alpha)
                ci += 1
            tree = merkle.MerkleProver.commit(_coord_columns(layer_eval))
            channel.mix_root(tree.root())
            folding_alpha = channel.draw_secure_felt()
            self.alphas.append(folding_alpha)
            self.inner.append((dom, layer_eval, tree))
            layer_eval = fold_line(layer_eval, dom, folding_alpha)
            dom = dom.double()
        assert ci == len(columns)
        # last layer
        coeffs = line_interpolate(layer_eval, dom)
        bound = 1 << config.log_last_layer_degree_bound
        assert all(c.is_zero() for c in coeffs[bound:]), "invalid degree"
        self.last_layer_poly = coeffs[:bound]
        self.last_layer_eval = layer_eval
        channel.mix_felts(self.last_layer_poly)

    def decommit(self, channel):
        max_log = self.columns[0][0]
        queries = generate_queries(channel, max_log, self.config.n_queries)
        by_size = query_positions_by_log_size(queries, max_log, [c[0] for c in self.columns])
        return self.decommit_on_queries(queries), by_size, queries

    def decommit_on_queries(self, queries) -> FriProof:
        max_log = self.columns[0][0]
        fri_witness = []
        pos_by_size = {}
        coord = []
        for log_c, col in self.columns:
            cq = fold_queries(queries, max_log - log_c)
            positions, w = _decommit_positions_and_witness(col, cq, 1)
            pos_by_size[log_c] = positions
            fri_witness.extend(w)
            coord.extend(_coord_columns(col))
        _, hw, cw = self.first_tree.decommit(pos_by_size, coord)
        first = FriLayerProof(fri_witness, MerkleDecommitment(hw, cw), self.first_tree.root())
        inner = []
        lq = fold_queries(queries, 1)
        for dom, ev, tree in self.inner:
            positions, w = _decommit_positions_and_witness(ev, lq, 1)
            _, hw, cw = tree.decommit({dom.log_size: positions}, _coord_columns(ev))
            inner.append(FriLayerProof(w, MerkleDecommitment(hw, cw), tree.root()))
            lq = fold_queries(lq, 1)
        return FriProof(first, inner, list(self.last_layer_poly), len(self.last_layer_poly).bit_length() - 1)


# ---- verifier ----------------------------------------------------------------
def _rebuild_evals(queries, query_evals, witness_iter, fold_step):
    """compute_decommitment_positions_and_rebuild_evals -> (positions, subsets[(start, [QM31..])])."""
    positions = []
    subsets = []
    qe = iter(query_evals)
    i = 0
    while i < len(queries):
        j = i
        while j < len(queries) and (queries[j] >> fold_step) == (queries[i] >> fold_step):
            j += 1
        subset = queries[i:j]
        start = (subset[0] >> fold_step) << fold_step
        k = 0
        evs = []
        for pos in range(start, start + (1 << fold_step)):
            positions.append(pos)
            if k < len(subset) and subset[k] == pos:
                k += 1
                evs.append(next(qe))
            else:
                try:
                    evs.append(next(witness_iter))
                except StopIteration:
                    raise FriVerificationError("InsufficientWitness")
        subsets.append((start, evs))
        i = j
    return positions, subsets


def _fold_pair_line(f_x: QM31, f_nx: QM31, x: int, alpha: QM31) -> QM31:
    xinv = pow(x, P - 2, P)
    return (f_x + f_nx) + alpha * ((f_x - f_nx) * xinv)


def _fold_pair_circle(f_p: QM31, f_np: QM31, y: int, alpha: QM31) -> QM31:
    yinv = pow(y, P - 2, P)
    return alpha * ((f_p - f_np) * yinv) + (f_p + f_np)


class FriVerifier:
    def __init__(self, channel, config, proof: FriProof, column_log_degree_bounds):
        """column_log_degree_bounds: descending, unique."""
        self.config = config
        self.proof = proof
        self.bounds = list(column_log_degree_bounds)
        channel.mix_root(proof.first_layer.commitment)
        self.first_alpha = channel.draw_secure_felt()
        self.column_logs = [b + config.log_blowup_factor for b in self.bounds]
        layer_bound = self.bounds[0] - 1
        layer_domain = LineDomain(Coset.half_odds(layer_bound + config.log_blowup_factor))
        self.inner = []
        for lp in proof.inner_layers:
            channel.mix_root(lp.commitment)
            alpha = channel.draw_secure_felt()
            self.inner.append((layer_bound, layer_domain, alpha, lp))
            layer_bound -= 1
            if layer_bound < 0:
                raise FriVerificationError("InvalidNumFriLayers")
            layer_domain = layer_domain.double()
        if layer_bound != config.log_last_layer_degree_bound:
            raise FriVerificationError("InvalidNumFriLayers")
        self.last_domain = layer_domain
        if len(proof.last_layer_poly) > (1 << config.log_last_layer_degree_bound):
            raise FriVerificationError("LastLayerDegreeInvalid")
        channel.mix_felts(proof.last_layer_poly)
        self.queries = None

    def sample_query_positions(self, channel):
        max_log = self.column_logs[0]
        self.queries = generate_queries(channel, max_log, self.config.n_queries)
        return query_positions_by_log_size(self.queries, max_log, self.column_logs)

    def decommit(self, first_layer_query_evals):
        """first_layer_query_evals[col] = [QM31 per (folded) query position]."""
        queries = self.queries
        max_log = self.column_logs[0]
        # first layer
        wit = iter(self.proof.first_layer.fri_witness)
        pos_by_size = {}
        decommitted = []
        sparse = []
        for log_c, evals in zip(self.column_logs, first_layer_query_evals):
            cq = fold_queries(queries, max_log - log_c)
            if len(cq) != len(evals):
                raise FriVerificationError("FirstLayerEvaluationsInvalid")
            positions, subsets = _rebuild_evals(cq, evals, wit, 1)
            pos_by_size[log_c] = positions
            sparse.append(subsets)
            decommitted.append([e for _, evs in subsets for e in evs])
        if next(wit, None) is not None:
            raise FriVerificationError("FirstLayerEvaluationsInvalid")
        col_sizes = [s for s in self.column_logs for _ in range(4)]
        # decommitted values: per layer (log size desc) per position per coordinate column
        flat = []
        for log_c, vals in zip(self.column_logs, decommitted):
            for v in vals:
                flat.extend(v.tup())
        d = self.proof.first_layer.decommitment
        try:
            merkle.verify(self.proof.first_layer.commitment, col_sizes, pos_by_size, flat, d.hash_witness, d.column_witness)
        except merkle.MerkleVerificationError as e:
            raise FriVerificationError(f"FirstLayerCommitmentInvalid: {e}")
        # inner layers
        layer_queries = fold_queries(queries, 1)
        layer_evals = [QM31() for _ in layer_queries]
        ci = 0
        # Circle columns are folded (and alpha^2-accumulated) with the *previous* layer's
        # folding alpha [artifact-verified: tests/test_oracle_kat.py].
        prev_alpha = self.first_alpha
        for bound, dom, alpha, lp in self.inner:
            while ci < len(self.bounds) and self.bounds[ci] - 1 == bound:
                cdom = CanonicCoset(self.column_logs[ci]).circle_domain()
                folded = []
                for start, evs in sparse[ci]:
                    p = cdom.at(bit_reverse_index(start, cdom.log_size))
                    folded.append(_fold_pair_circle(evs[0], evs[1], p[1], prev_alpha))
                a2 = prev_alpha * prev_alpha
                assert len(folded) == len(layer_evals)
                layer_evals = [le * a2 + f for le, f in zip(layer_evals, folded)]
                ci += 1
            wit = iter(lp.fri_witness)
            positions, subsets = _rebuild_evals(layer_queries, layer_evals, wit, 1)
            if next(wit, None) is not None:
                raise FriVerificationError("InnerLayerEvaluationsInvalid")
            flat = []
            for _, evs in subsets:
                for v in evs:
                    flat.extend(v.tup())
            try:
                merkle.verify(lp.commitment, [dom.log_size] * 4, {dom.log_size: positions}, flat,
                              lp.decommitment.hash_witness, lp.decommitment.column_witness)
            except merkle.MerkleVerificationError as e:
                raise FriVerificationError(f"InnerLayerCommitmentInvalid: {e}")
            new_evals = []
            for start, evs in subsets:
                x = dom.at(bit_reverse_index(start, dom.log_size))
                new_evals.append(_fold_pair_line(evs[0], evs[1], x, alpha))
            layer_queries = fold_queries(layer_queries, 1)
            layer_evals = new_evals
            prev_alpha = alpha
        if ci != len(self.bounds):
            raise FriVerificationError("columns left unfolded")
        # last layer
        for q, ev in zip(layer_queries, layer_evals):
            x = self.last_domain.at(bit_reverse_index(q, self.last_domain.log_size))
            if ev != line_poly_eval(self.proof.last_layer_poly, QM31(x, 0, 0, 0)):
                raise FriVerificationError("LastLayerEvaluationsInvalid")
