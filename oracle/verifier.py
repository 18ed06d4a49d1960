"""STARK verifier (oracle; test infrastructure only).

Restates ``stwo::core::verifier::verify`` + ``CommitmentSchemeVerifier`` and the
LuminAIR wrapper crates/verifiers/rust/src/verifier.rs:21-143.  This is the
acceptance oracle: a proof produced by the CUDA prover must pass it, and it
must accept the proof the reference commits (ui/demo/public/proof).
"""
from __future__ import annotations

from . import merkle
from .air import (
    INTERACTION_TRACE_IDX,
    ORIGINAL_TRACE_IDX,
    PREPROCESSED_TRACE_IDX,
    AddEval,
    ContiguousEval,
    Exp2Eval,
    Exp2LookupEval,
    FrameworkComponent,
    LessThanEval,
    Log2Eval,
    Log2LookupEval,
    MaxReduceEval,
    RangeCheckLookupEval,
    RecipEval,
    RemEval,
    SinEval,
    SinLookupEval,
    SqrtEval,
    SumReduceEval,
    InputsEval,
    MulEval,
    PointEvaluationAccumulator,
    RelationElements,
    TraceLocationAllocator,
)
from .channel import Blake2sChannel
from .circle import CanonicCoset, bit_reverse_index
from .fields import QM31
from .fri import FriVerificationError, FriVerifier
from .proof import LuminairProof, StarkProof
from .quotients import accumulate_row_quotients, column_sample_batches, quotient_constants


class VerificationError(Exception):
    pass


def get_random_point(channel):
    """CirclePoint::<SecureField>::get_random_point."""
    t = channel.draw_secure_felt()
    t2 = t.square()
    inv = (t2 + 1).inv()
    x = (QM31(1, 0, 0, 0) - t2) * inv
    y = (t + t) * inv
    return (x, y)


class CommitmentSchemeVerifier:
    def __init__(self, config):
        self.config = config
        self.trees = []  # (root, extended column log sizes)

    def commit(self, root: bytes, log_sizes, channel):
        channel.mix_root(root)
        self.trees.append((root, [s + self.config.log_blowup_factor for s in log_sizes]))

    def column_log_sizes(self):
        return [t[1] for t in self.trees]

    def verify_values(self, sampled_points, proof: StarkProof, channel):
        flat_vals = [v for tree in proof.sampled_values for col in tree for v in col]
        channel.mix_felts(flat_vals)
        random_coeff = channel.draw_secure_felt()
        all_sizes = sorted({s for t in self.column_log_sizes() for s in t}, reverse=True)
        bounds = [s - self.config.log_blowup_factor for s in all_sizes]
        fri = FriVerifier(channel, self.config, proof.fri_proof, bounds)
        channel.mix_u64(proof.proof_of_work)
        if channel.trailing_zeros() < self.config.pow_bits:
            raise VerificationError("ProofOfWork")
        qpos = fri.sample_query_positions(channel)
        for (root, sizes), dec, qv in zip(self.trees, proof.decommitments, proof.queried_values):
            try:
                merkle.verify(root, sizes, qpos, qv, dec.hash_witness, dec.column_witness)
            except merkle.MerkleVerificationError as e:
                raise VerificationError(f"Merkle: {e}")
        # fri answers
        samples = []  # flat per column: [(point, value)]
        sizes_flat = []
        for t, (root, sizes) in enumerate(self.trees):
            for c, s in enumerate(sizes):
                pts = sampled_points[t][c]
                vals = proof.sampled_values[t][c]
                if len(pts) != len(vals):
                    raise VerificationError("InvalidStructure: sample count")
                samples.append(list(zip(pts, vals)))
                sizes_flat.append(s)
        qv_iters = [iter(q) for q in proof.queried_values]
        answers = []
        for log_size in all_sizes:
            idx = [i for i, s in enumerate(sizes_flat) if s == log_size]
            n_cols_per_tree = [sum(1 for s in sizes if s == log_size) for _, sizes in self.trees]
            batches = column_sample_batches([samples[i] for i in idx])
            consts = quotient_constants(batches, random_coeff)
            dom = CanonicCoset(log_size).circle_domain()
            evs = []
            for q in qpos[log_size]:
                px, py = dom.at(bit_reverse_index(q, log_size))
                row = []
                for it, n in zip(qv_iters, n_cols_per_tree):
                    row.extend(next(it) for _ in range(n))
                evs.append(accumulate_row_quotients(batches, consts, row, px, py))
            answers.append(evs)
        try:
            fri.decommit(answers)
        except FriVerificationError as e:
            raise VerificationError(f"Fri: {e}")
        return fri


def stark_verify(components, channel, scheme: CommitmentSchemeVerifier, proof: StarkProof, _skip_oods=False):
    """stwo::core::verifier::verify."""
    random_coeff = channel.draw_secure_felt()
    comp_log = max(c.max_constraint_log_degree_bound() for c in components)
    scheme.commit(proof.commitments[-1], [comp_log] * 4, channel)
    oods = get_random_point(channel)
    n_trees = len(scheme.trees)
    sample_points = [[] for _ in range(n_trees)]
    # preprocessed columns: only those referenced are sampled at the point
    n_pre = len(scheme.trees[PREPROCESSED_TRACE_IDX][1])
    pre_pts = [[] for _ in range(n_pre)]
    for c in components:
        mp = c.mask_points(oods)
        for i in c.preprocessed_indices:
            if not pre_pts[i]:
                pre_pts[i] = [oods]
        for t in (ORIGINAL_TRACE_IDX, INTERACTION_TRACE_IDX):
            sample_points[t].extend(mp[t])
    sample_points[PREPROCESSED_TRACE_IDX] = pre_pts
    sample_points[n_trees - 1] = [[oods] for _ in range(4)]
    sv = proof.sampled_values
    comp_mask = sv[-1]
    if len(comp_mask) != 4 or any(len(c) != 1 for c in comp_mask):
        raise VerificationError("InvalidStructure: composition mask")
    composition_oods = QM31.from_partial_evals([c[0] for c in comp_mask])
    acc = PointEvaluationAccumulator(random_coeff)
    for c in components:
        c.evaluate_constraint_quotients_at_point(oods, sv, acc)
    if composition_oods != acc.finalize() and not _skip_oods:
        raise VerificationError("OodsNotMatching")
    return scheme.verify_values(sample_points, proof, channel)


# ---------------------------------------------------------------------------
# LuminAIR wrapper (crates/verifiers/rust/src/verifier.rs)
# ---------------------------------------------------------------------------
# component slot order: crates/air/src/lib.rs:30-48.  The UI artifact predates
# the 17-slot schema and has 8 slots with add, mul first.
SLOT_EVALS = {0: AddEval, 1: MulEval, 2: RecipEval, 3: SinEval, 4: SinLookupEval, 5: SumReduceEval, 6: MaxReduceEval,
              7: SqrtEval, 8: RemEval, 9: Exp2Eval, 10: Exp2LookupEval, 11: Log2Eval, 12: Log2LookupEval,
              13: LessThanEval, 14: RangeCheckLookupEval, 15: InputsEval, 16: ContiguousEval}


def lut_log_sizes(preprocessed):
    """{lut name: log_size} from the preprocessed column list [(id, log_size)]
    (``lookups.<lut>.layout.log_size``, crates/air/src/components/mod.rs:299,366,399,432)."""
    out = {}
    for cid, log in preprocessed:
        for name in ("sin", "exp2", "log2", "range_check"):
            if cid.startswith(name + "_"):
                out[name] = log
    return out


def sort_preprocessed(columns):
    """PreProcessedTrace::new (preprocessed.rs:152-155): stable sort by log_size, descending.
    columns: [(id, log_size or values, ...)] with the log size derivable from item[1]."""
    def log_of(item):
        v = item[1]
        return v if isinstance(v, int) else (len(v).bit_length() - 1)
    return sorted(columns, key=lambda it: -log_of(it))


def luminair_components(claim, interaction_claim, node_elements, slot_evals=SLOT_EVALS, preprocessed=(), lookups=None):
    """LuminairComponents::new (crates/air/src/components/mod.rs:261-527).  preprocessed: [(id, log_size)] in
    committed (sorted) order."""
    alloc = TraceLocationAllocator([cid for cid, _ in preprocessed])
    lut_logs = lut_log_sizes(preprocessed)
    comps = []
    for slot, log_size in enumerate(claim):
        if log_size is None:
            continue
        cls = slot_evals[slot]
        comps.append(FrameworkComponent(alloc, cls(log_size, node_elements, lookups, lut_logs), interaction_claim[slot]))
    return comps


def log_sum_valid(interaction_claim) -> bool:
    """crates/air/src/utils.rs:29-57: sum of claimed sums == 0."""
    s = QM31()
    for c in interaction_claim:
        if c is not None:
            s = s + c
    return s.is_zero()


def draw_interaction_elements(channel):
    """LuminairInteractionElements::draw, crates/air/src/components/mod.rs:227-235 +
    lookups/mod.rs:44-51: NodeElements(2), then sin(2), exp2(2), log2(2), range_check(1)."""
    node = RelationElements.draw(channel, 2)
    lookups = {
        "sin": RelationElements.draw(channel, 2),
        "exp2": RelationElements.draw(channel, 2),
        "log2": RelationElements.draw(channel, 2),
        "range_check": RelationElements.draw(channel, 1),
    }
    return node, lookups


def verify(proof: LuminairProof, channel_variant="legacy", slot_evals=SLOT_EVALS, draw_lookup_elements=True,
           preprocessed=(), config=None, _skip_oods=False):
    """crates/verifiers/rust/src/verifier.rs:21-143.  Returns the channel (for transcript tests).
    preprocessed: [(id, log_size)] of the LUT columns named by the circuit settings (any order)."""
    sp = proof.proof
    config = config or sp.config
    channel = Blake2sChannel(channel_variant)
    scheme = CommitmentSchemeVerifier(config)
    # preprocessed
    preprocessed = sort_preprocessed(list(preprocessed))
    scheme.commit(sp.commitments[PREPROCESSED_TRACE_IDX], [log for _, log in preprocessed], channel)
    # main: claim.mix_into then commit
    main_sizes, int_sizes = [], []
    for slot, log_size in enumerate(proof.claim):
        if log_size is None:
            continue
        channel.mix_u64(log_size)
        cls = slot_evals[slot]
        main_sizes += [log_size] * cls.n_main
        int_sizes += [log_size] * (4 * cls.n_interaction)
    scheme.commit(sp.commitments[ORIGINAL_TRACE_IDX], main_sizes, channel)
    lookups = None
    if draw_lookup_elements:
        node, lookups = draw_interaction_elements(channel)
    else:
        node = RelationElements.draw(channel, 2)
    if not log_sum_valid(proof.interaction_claim):
        raise VerificationError("InvalidLogupSum")
    for c in proof.interaction_claim:
        if c is not None:
            channel.mix_felts([c])
    scheme.commit(sp.commitments[INTERACTION_TRACE_IDX], int_sizes, channel)
    comps = luminair_components(proof.claim, proof.interaction_claim, node, slot_evals, preprocessed, lookups)
    stark_verify(comps, channel, scheme, sp, _skip_oods)
    return channel
