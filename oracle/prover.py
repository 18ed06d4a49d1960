"""STARK prover (oracle; test infrastructure only).

CPU restatement of ``CommitmentSchemeProver`` / ``TreeBuilder`` / ``prover::prove``
(stwo @0790eba, un-vendored) and of LuminAIR's orchestration
crates/prover/src/prover.rs:28-319.  The CUDA path (luminair_b200) must
reproduce this module's outputs bit-for-bit; this module in turn reproduces the
reference's committed proof byte-for-byte (tests/test_oracle_prover.py).
"""
from __future__ import annotations

import numpy as np

from . import cfft, merkle
from .air import (
    INTERACTION_TRACE_IDX,
    ORIGINAL_TRACE_IDX,
    PREPROCESSED_TRACE_IDX,
    PointEvaluationAccumulator,
    RelationElements,
    gen_interaction_trace,
)
from .channel import Blake2sChannel, grind
from .circle import CanonicCoset, index_to_point, pt_add
from .fields import P, U64, QM31, m_add
from .fri import FriProver
from .proof import LuminairProof, MerkleDecommitment, PcsConfig, StarkProof
from .quotients import compute_fri_quotients
from .verifier import SLOT_EVALS, draw_interaction_elements, get_random_point, luminair_components, sort_preprocessed


class ProvingError(Exception):
    pass


class CommitmentTreeProver:
    def __init__(self, polys, log_blowup, channel):
        """polys: list of coefficient arrays (uint64).  Evaluates each on
        CanonicCoset(log + blowup), Merkle-commits, mixes the root."""
        self.polys = polys
        self.evals = []
        for c in polys:
            log = len(c).bit_length() - 1
            dom = CanonicCoset(log + log_blowup).circle_domain()
            self.evals.append(cfft.evaluate(c, dom))
        self.tree = merkle.MerkleProver.commit(self.evals)
        channel.mix_root(self.tree.root())

    def decommit(self, queries_per_log_size):
        qv, hw, cw = self.tree.decommit(queries_per_log_size, self.evals)
        return qv, MerkleDecommitment(hw, cw)


class TreeBuilder:
    def __init__(self, scheme):
        self.scheme = scheme
        self.polys = []

    def extend_evals(self, columns):
        """columns: evaluations on CanonicCoset(log).circle_domain(), storage order."""
        for col in columns:
            col = np.asarray(col, dtype=U64)
            log = len(col).bit_length() - 1
            self.polys.append(cfft.interpolate(col, CanonicCoset(log).circle_domain()))

    def extend_polys(self, polys):
        self.polys.extend(polys)

    def commit(self, channel):
        self.scheme.trees.append(CommitmentTreeProver(self.polys, self.scheme.config.log_blowup_factor, channel))


class CommitmentSchemeProver:
    def __init__(self, config: PcsConfig):
        self.config = config
        self.trees = []

    def tree_builder(self):
        return TreeBuilder(self)

    def roots(self):
        return [t.tree.root() for t in self.trees]

    def prove_values(self, sampled_points, channel):
        """sampled_points[tree][col] = [points].  -> dict of proof parts."""
        sampled_values = []
        samples_flat = []
        for t, tree in enumerate(self.trees):
            tv = []
            for c, poly in enumerate(tree.polys):
                vals = [cfft.eval_at_point(poly, pt[0], pt[1]) for pt in sampled_points[t][c]]
                tv.append(vals)
                samples_flat.append(list(zip(sampled_points[t][c], vals)))
            sampled_values.append(tv)
        channel.mix_felts([v for tv in sampled_values for col in tv for v in col])
        random_coeff = channel.draw_secure_felt()
        columns = [e for tree in self.trees for e in tree.evals]
        quotients = compute_fri_quotients(columns, samples_flat, random_coeff)
        fri = FriProver(channel, self.config, quotients)
        nonce = grind(channel, self.config.pow_bits)
        channel.mix_u64(nonce)
        fri_proof, qpos, _ = fri.decommit(channel)
        decommitments, queried = [], []
        for tree in self.trees:
            qv, dec = tree.decommit(qpos)
            queried.append(qv)
            decommitments.append(dec)
        self.debug = {"quotients": quotients, "fri": fri, "random_coeff": random_coeff}
        return StarkProof(self.config, self.roots(), sampled_values, decommitments, queried, nonce, fri_proof)


def compute_composition_polynomial(components, random_coeff: QM31, lde_trace, polys=None):
    """ComponentProvers::compute_composition_polynomial + DomainEvaluationAccumulator::finalize.
    lde_trace[tree] = list of LDE columns.  -> 4 coefficient arrays of log = max bound."""
    total = sum(c.n_constraints for c in components)
    powers = [QM31(1, 0, 0, 0)]
    for _ in range(total - 1):
        powers.append(powers[-1] * random_coeff)
    max_log = max(c.max_constraint_log_degree_bound() for c in components)
    sub = {}
    for c in components:
        n = c.n_constraints
        mine = powers[len(powers) - n:]
        powers = powers[: len(powers) - n]
        mine = mine[::-1]
        acc = c.evaluate_constraint_quotients_on_domain(lde_trace, mine, polys)
        log = c.max_constraint_log_degree_bound()
        sub[log] = acc if log not in sub else sub[log] + acc
    assert not powers
    cur = None
    for log in sorted(sub):
        vals = sub[log]
        dom = CanonicCoset(log).circle_domain()
        if cur is not None:
            lifted = QM31(*[cfft.evaluate(c, dom) for c in cur])
            vals = vals + lifted
        cur = [cfft.interpolate(np.asarray(c, dtype=U64), dom) for c in vals.c]
    if len(cur[0]) != (1 << max_log):
        raise AssertionError("composition log size")
    return cur


def stark_prove(components, channel, scheme: CommitmentSchemeProver) -> StarkProof:
    """stwo::prover::prove."""
    random_coeff = channel.draw_secure_felt()
    lde = {t: scheme.trees[t].evals for t in (PREPROCESSED_TRACE_IDX, ORIGINAL_TRACE_IDX, INTERACTION_TRACE_IDX)}
    polys = {t: scheme.trees[t].polys for t in (PREPROCESSED_TRACE_IDX, ORIGINAL_TRACE_IDX, INTERACTION_TRACE_IDX)}
    comp = compute_composition_polynomial(components, random_coeff, lde, polys)
    tb = scheme.tree_builder()
    tb.extend_polys(comp)
    tb.commit(channel)
    oods = get_random_point(channel)
    n_trees = len(scheme.trees)
    sample_points = [[] for _ in range(n_trees)]
    pre_pts = [[] for _ in scheme.trees[PREPROCESSED_TRACE_IDX].polys]
    for c in components:
        mp = c.mask_points(oods)
        for i in c.preprocessed_indices:
            if not pre_pts[i]:
                pre_pts[i] = [oods]
        for t in (ORIGINAL_TRACE_IDX, INTERACTION_TRACE_IDX):
            sample_points[t].extend(mp[t])
    sample_points[PREPROCESSED_TRACE_IDX] = pre_pts
    sample_points[n_trees - 1] = [[oods] for _ in range(4)]
    proof = scheme.prove_values(sample_points, channel)
    composition_oods = QM31.from_partial_evals([c[0] for c in proof.sampled_values[-1]])
    acc = PointEvaluationAccumulator(random_coeff)
    for c in components:
        c.evaluate_constraint_quotients_at_point(oods, proof.sampled_values, acc)
    if composition_oods != acc.finalize():
        raise ProvingError("ConstraintsNotSatisfied")
    return proof


# ---------------------------------------------------------------------------
# LuminAIR prove()  (crates/prover/src/prover.rs:28-319)
# ---------------------------------------------------------------------------
SLOT_OF = {"add": 0, "mul": 1, "recip": 2, "sin": 3, "sin_lookup": 4, "sum_reduce": 5, "max_reduce": 6, "sqrt": 7, "rem": 8,
           "exp2": 9, "exp2_lookup": 10, "log2": 11, "log2_lookup": 12, "less_than": 13, "range_check_lookup": 14,
           "inputs": 15, "contiguous": 16}  # LuminairClaim field order, crates/air/src/lib.rs:30-48
N_LANES = 16


def pad_table(rows: np.ndarray, padding_row) -> np.ndarray:
    """write_trace: pad to max(next_pow2(n), 16) with the component's padding row
    (e.g. add/witness.rs:43-46).  rows: [n, n_cols]."""
    n = rows.shape[0]
    if n == 0:
        raise ValueError("TraceError::EmptyTrace")
    size = max(1 << (n - 1).bit_length(), N_LANES)
    out = np.tile(np.asarray(padding_row, dtype=U64), (size, 1))
    out[:n] = rows
    return out


def prove(pie, n_slots: int = 17, slot_evals=SLOT_EVALS, slot_of=SLOT_OF, channel_variant="legacy",
          config: PcsConfig | None = None, draw_lookup_elements=True, return_debug=False, preprocessed=()):
    """pie: list of (name, rows[n, n_cols] M31 values) in ``pie.trace_tables`` order.
    preprocessed: [(id, values[2^k])] LUT columns in ``lookups_to_preprocessed_column`` order
    (preprocessed.rs:181-206: sin 0/1, exp2 0/1, log2 0/1, range_check 0); they are generated by the caller
    (host libm, preprocessed.rs:351-383) - the prover only commits to them."""
    config = config or PcsConfig()
    channel = Blake2sChannel(channel_variant)
    scheme = CommitmentSchemeProver(config)
    # phase 0: preprocessed (prover.rs:52-59)
    preprocessed = sort_preprocessed([(cid, np.asarray(v, dtype=U64)) for cid, v in preprocessed])
    pre_vals = dict(preprocessed)
    pre_meta = [(cid, len(v).bit_length() - 1) for cid, v in preprocessed]
    tb = scheme.tree_builder()
    tb.extend_evals([v for _, v in preprocessed])
    tb.commit(channel)
    # phase 1: main trace
    tb = scheme.tree_builder()
    claim = [None] * n_slots
    mains = {}
    for name, rows in pie:
        slot = slot_of[name]
        cls = slot_evals[slot]
        padded = pad_table(np.asarray(rows, dtype=U64), cls.padding_row())
        log_size = padded.shape[0].bit_length() - 1
        cols = [np.ascontiguousarray(padded[:, j]) for j in range(padded.shape[1])]
        assert len(cols) == cls.n_main
        tb.extend_evals(cols)
        claim[slot] = log_size
        mains[slot] = cols
    for c in claim:
        if c is not None:
            channel.mix_u64(c)
    tb.commit(channel)
    # phase 2: interaction trace
    lookups = None
    if draw_lookup_elements:
        node, lookups = draw_interaction_elements(channel)
    else:
        node = RelationElements.draw(channel, 2)
    tb = scheme.tree_builder()
    iclaim = [None] * n_slots
    for slot in range(n_slots):
        if claim[slot] is None:
            continue
        cols, cs = gen_interaction_trace(slot_evals[slot], mains[slot], claim[slot], node, lookups, pre_vals)
        tb.extend_evals(cols)
        iclaim[slot] = cs
    for c in iclaim:
        if c is not None:
            channel.mix_felts([c])
    tb.commit(channel)
    comps = luminair_components(claim, iclaim, node, slot_evals, pre_meta, lookups)
    proof = stark_prove(comps, channel, scheme)
    lp = LuminairProof(claim, iclaim, proof)
    if return_debug:
        return lp, {"scheme": scheme, "channel": channel, "components": comps, "node": node, "lookups": lookups}
    return lp
