"""prove() composed from the backend-TRAIT entry points of the C ABI - the call sequence a Rust ``CudaBackend`` shim produces
when stwo's generic code drives it (path B), with the Fiat-Shamir channel and all bookkeeping on the caller's
side exactly where stwo keeps them.

It replays /root/reference/crates/prover/src/prover.rs:38-312 one trait method at a time:

    precompute_twiddles            lb_twiddles_ensure                     prover.rs:38-42
    extend_evals -> interpolate    lb_interpolate_batch                   prover.rs:57, */witness.rs (add/witness.rs:51,164)
    commit -> evaluate + Merkle    lb_evaluate_batch, lb_merkle_commit_layer   prover.rs:59,179,298
    LogupTraceGenerator            lb_logup_interaction_trace_lut         */witness.rs (add/witness.rs:126-167)
    constraint quotients           lb_constraint_quotients_lut            components/mod.rs:530-601
    accumulate / finalize          lb_accumulate, lb_evaluate_batch, lb_interpolate_batch
    eval_at_point                  lb_eval_at_point
    accumulate_quotients           lb_accumulate_quotients
    fold_circle_into_line / line   lb_fold_circle_into_line, lb_fold_line
    grind                          lb_grind
    decommitment reads             lb_download (single words)

and must return the same bytes as ``lb_prove`` (tests/test_gpu_traits_prover.py): evidence that the trait-level surface is
complete and composes, which is all a binding without a Rust toolchain can show.  Host-side scalars: ``hostmath``.
"""
from __future__ import annotations

import numpy as np

from . import proof as wire
from .backend import ColumnBatch, CudaBackend
from .hostmath import ONE, P, QM31, Blake2sChannel, bit_reverse, blake2s, index_to_point, m_inv, qpt_add, subgroup_gen
from .prover import CLAIM_SLOT, PcsConfig, ProvingError, _lut_column

# LB_COMP_* ids (include/luminair_b200.h) and the shape of every component: main columns, LogUp fractions, constraints, the LUT
# relation it uses (LB_REL_*), preprocessed columns it tabulates, non-zero entries of its padding row (<name>/table.rs padding())
COMPONENTS = {
    "add": (0, 15, 3, 9, 0, 0, {4: 1}), "mul": (1, 16, 3, 9, 0, 0, {4: 1}), "inputs": (2, 7, 1, 4, 0, 0, {2: 1}),
    "sum_reduce": (4, 14, 2, 9, 0, 0, {3: 1}), "max_reduce": (5, 15, 2, 11, 0, 0, {3: 1}),
    "contiguous": (6, 11, 2, 6, 0, 0, {3: 1}), "recip": (7, 13, 2, 7, 0, 0, {3: 1}), "sqrt": (8, 13, 2, 7, 0, 0, {3: 1}),
    "rem": (9, 16, 3, 9, 0, 0, {4: 1}), "sin": (10, 12, 3, 7, 1, 0, {3: 1}), "exp2": (11, 12, 3, 7, 2, 0, {3: 1}),
    "log2": (12, 12, 3, 7, 3, 0, {3: 1}), "sin_lookup": (13, 1, 1, 1, 1, 2, {}), "exp2_lookup": (14, 1, 1, 1, 2, 2, {}),
    "log2_lookup": (15, 1, 1, 1, 3, 2, {}), "less_than": (16, 22, 7, 16, 4, 0, {4: 1, 10: 1, 11: 4096, 12: 1, 14: 1}),
    "range_check_lookup": (17, 1, 1, 1, 4, 1, {}),
}


class _Tree:
    """One commitment tree: coefficient columns, their LDEs, the Merkle layers (all device-resident)."""

    def __init__(self):
        self.polys = []   # (ptr, log)
        self.ldes = []    # (ptr, log)
        self.layers = {}  # log -> device buffer of 2^log digests
        self.max_log = -1
        self.root = None
        self.keep = []

    def sorted_cols(self):
        return sorted(self.ldes, key=lambda c: -c[1])  # stable: by size, descending


def _merkle_commit(be: CudaBackend, tree: _Tree):
    """MerkleProver::commit through MerkleOps::commit_on_layer, one layer per call."""
    if not tree.ldes:
        tree.root = blake2s(b"")
        return
    cols = tree.sorted_cols()
    tree.max_log = cols[0][1]
    prev = None
    for log in range(tree.max_log, -1, -1):
        out = be.alloc(8 << log)
        be.merkle_commit_layer(log, prev.ptr if prev is not None else None, [p for p, l in cols if l == log], out.ptr)
        tree.layers[log] = out
        prev = out
    tree.root = be.download(tree.layers[0], 8).tobytes()


def _commit(be, tree: _Tree, runs, blowup: int, channel: Blake2sChannel):
    """TreeBuilder::commit: evaluate every polynomial on CanonicCoset(log + blowup), Merkle-commit, mix the root."""
    for buf, n_cols, log in runs:
        lde = be.alloc(n_cols << (log + blowup))
        be.evaluate(ColumnBatch(buf, n_cols, log), ColumnBatch(lde, n_cols, log + blowup))
        tree.keep += [buf, lde]
        for k in range(n_cols):
            tree.polys.append((buf.at(k << log), log))
            tree.ldes.append((lde.at(k << (log + blowup)), log + blowup))
    _merkle_commit(be, tree)
    channel.mix_root(tree.root)


def _word(be, ptr: int, idx: int) -> int:
    return int(be.download_ptr(ptr + 4 * idx, 1)[0])


def _decommit(be, tree: _Tree, queries_per_log):
    """MerkleProver::decommit -> (queried values, Decommitment)."""
    qv, dec = [], wire.Decommitment()
    if tree.max_log < 0:
        return qv, dec
    cols = tree.sorted_cols()
    last = []
    for log in range(tree.max_log, -1, -1):
        layer_cols = [p for p, l in cols if l == log]
        prev = tree.layers.get(log + 1)
        col_q = list(queries_per_log.get(log, []))
        pi = ci = 0
        total = []
        while pi < len(last) or ci < len(col_q):
            cands = ([last[pi] // 2] if pi < len(last) else []) + ([col_q[ci]] if ci < len(col_q) else [])
            node = min(cands)
            if prev is not None:
                for child in (2 * node, 2 * node + 1):
                    if pi < len(last) and last[pi] == child:
                        pi += 1
                    else:
                        dec.hash_witness.append(be.download_ptr(prev.ptr + 32 * child, 8).tobytes())
            vals = [_word(be, p, node) for p in layer_cols]
            if ci < len(col_q) and col_q[ci] == node:
                ci += 1
                qv.extend(vals)
            else:
                dec.column_witness.extend(vals)
            total.append(node)
        last = total
    return qv, dec


def _fold_queries(qs, n):
    out = []
    for q in qs:
        f = q >> n
        if not out or out[-1] != f:
            out.append(f)
    return out


def _fri_positions_and_witness(be, coord_ptrs, queries):
    positions, witness = [], []
    i = 0
    while i < len(queries):
        j = i
        while j < len(queries) and (queries[j] >> 1) == (queries[i] >> 1):
            j += 1
        start, k = (queries[i] >> 1) << 1, i
        for pos in (start, start + 1):
            positions.append(pos)
            if k < j and queries[k] == pos:
                k += 1
                continue
            witness.append(tuple(_word(be, p, pos) for p in coord_ptrs))
        i = j
    return positions, witness


def _coset_vanishing(log: int, pt):
    """core/constraints.rs coset_vanishing of CanonicCoset(log).coset at a QM31 point."""
    a = index_to_point(-subgroup_gen(log + 1))
    b = index_to_point(subgroup_gen(log) >> 1)
    q = qpt_add(qpt_add(pt, (QM31(a[0]), QM31(a[1]))), (QM31(b[0]), QM31(b[1])))
    x = q[0]
    for _ in range(1, log):
        x = x * x * 2 - ONE
    return x


def prove_with_traits(pie, backend: CudaBackend, preprocessed=(), config: PcsConfig | None = None,
                      channel_variant: str = "legacy") -> bytes:
    """pie / preprocessed as for ``prover.prove``.  Every O(N) step is one backend-trait call; the transcript lives here."""
    be = backend
    cfg = config or PcsConfig()
    blowup = cfg.log_blowup_factor
    ch = Blake2sChannel(channel_variant)
    trees = [_Tree() for _ in range(4)]
    n_slots = 17

    # ---- sizes, twiddles (prover.rs:38-42)
    tables = []
    for name, rows in pie:
        rows = np.ascontiguousarray(np.asarray(rows), dtype=np.uint32)
        kind, n_main, n_fracs, n_cons, lut, n_pre, pad = COMPONENTS[name]
        if rows.shape[0] == 0:
            raise ValueError("TraceError::EmptyTrace")
        log = max((rows.shape[0] - 1).bit_length(), 4)
        padded = np.zeros((1 << log, n_main), dtype=np.uint32)
        for c, v in pad.items():
            padded[:, c] = v
        padded[: rows.shape[0]] = rows
        tables.append((name, log, np.ascontiguousarray(padded.T)))
    pre_cols = []
    for cid, values in preprocessed:
        v = np.ascontiguousarray(values, dtype=np.uint32).reshape(-1)
        pre_cols.append((_lut_column(cid), v.size.bit_length() - 1, v))
    order = sorted(range(len(pre_cols)), key=lambda i: -pre_cols[i][1])  # PreProcessedTrace::new: by size, descending (stable)
    pre_cols = [pre_cols[i] for i in order]
    max_log = max([t[1] for t in tables] + [p[1] for p in pre_cols])
    be.precompute_twiddles(max_log + 1 + blowup)

    # ---- phase 0: preprocessed trace (prover.rs:52-59)
    runs, pre_vals, lut_log = [], [], {}
    for (lut, col), log, v in pre_cols:
        vals = be.upload(v)
        coeffs = be.upload(v)
        be.interpolate(ColumnBatch(coeffs, 1, log))
        runs.append((coeffs, 1, log))
        pre_vals.append(((lut, col), vals, log))
        lut_log[lut] = log
    _commit(be, trees[0], runs, blowup, ch)

    # ---- phase 1: main trace (prover.rs:61-179)
    comps, claim, runs = {}, [None] * n_slots, []
    main_at = 0
    for name, log, cols in tables:
        kind, n_main, n_fracs, n_cons, lut, n_pre, _ = COMPONENTS[name]
        vals = be.upload(cols.reshape(-1))
        coeffs = be.upload(cols.reshape(-1))
        be.interpolate(ColumnBatch(coeffs, n_main, log))
        runs.append((coeffs, n_main, log))
        claim[CLAIM_SLOT[name]] = log
        consumer = lut and not n_pre
        comps[CLAIM_SLOT[name]] = dict(name=name, kind=kind, log=log, vals=vals, n_main=n_main, n_fracs=n_fracs, n_cons=n_cons,
                                       lut=lut, n_pre=n_pre, eval_log=(max(log, lut_log[lut]) if consumer else log) + 1)
    for c in claim:
        if c is not None:
            ch.mix_u64(c)
    _commit(be, trees[1], runs, blowup, ch)

    # ---- phase 2: interaction trace (prover.rs:181-298)
    rels = []
    for _ in range(5):  # LuminairInteractionElements::draw: node, sin, exp2, log2, range_check
        z, alpha = ch.draw_secure_felts(2)
        rels.append((z.words(), alpha.words()))
    runs, inter_at = [], 0
    order = sorted(comps)  # claim-slot order = LuminairComponents order
    for slot in order:
        c = comps[slot]
        n_ic = 4 * c["n_fracs"]
        inter = be.alloc(n_ic << c["log"])
        lut_ptrs = [v.ptr for (l, col), v, _ in sorted([p for p in pre_vals if p[0][0] == c["lut"]], key=lambda p: p[0][1])] \
            if c["n_pre"] else []
        c["pre_idx"] = [i for i, p in enumerate(pre_vals) if c["n_pre"] and p[0][0] == c["lut"]]
        c["pre_idx"].sort(key=lambda i: pre_vals[i][0][1])
        claimed = be.logup_interaction_trace_lut(c["kind"], ColumnBatch(c["vals"], c["n_main"], c["log"]),
                                                 ColumnBatch(inter, n_ic, c["log"]), rels, lut_ptrs)
        c["claimed"] = QM31(*[int(x) for x in claimed])
        be.interpolate(ColumnBatch(inter, n_ic, c["log"]))
        runs.append((inter, n_ic, c["log"]))
        c["main_loc"], c["inter_loc"] = main_at, inter_at
        main_at += c["n_main"]
        inter_at += n_ic
    for slot in order:
        ch.mix_felts([comps[slot]["claimed"]])
    _commit(be, trees[2], runs, blowup, ch)

    # ---- stwo::prover::prove: composition polynomial
    random_coeff = ch.draw_secure_felt()
    total = sum(comps[s]["n_cons"] for s in order)
    powers = be.generate_secure_powers(random_coeff.words(), total)
    acc, remaining, keep = {}, total, []
    for slot in order:
        c = comps[slot]
        ev = c["eval_log"]

        def on_eval_domain(tree, first, n_cols):
            ptr, log = tree.ldes[first]
            if log == ev:
                return ColumnBatch(_Raw(ptr), n_cols, ev)
            cptr, clog = tree.polys[first]  # constraint-framework need_to_extend: re-evaluate on the evaluation domain
            ext = be.alloc(n_cols << ev)
            keep.append(ext)
            be.evaluate(ColumnBatch(_Raw(cptr), n_cols, clog), ColumnBatch(ext, n_cols, ev))
            return ColumnBatch(ext, n_cols, ev)

        main_ev = on_eval_domain(trees[1], c["main_loc"], c["n_main"])
        inter_ev = on_eval_domain(trees[2], c["inter_loc"], 4 * c["n_fracs"])
        lut_ptrs = [on_eval_domain(trees[0], i, 1).ptr for i in c["pre_idx"]]
        fresh = ev not in acc
        if fresh:
            acc[ev] = be.alloc(4 << ev)
        mine = [powers[remaining - 1 - k] for k in range(c["n_cons"])]  # first constraint takes the highest power
        remaining -= c["n_cons"]
        be.constraint_quotients_lut(c["kind"], main_ev, inter_ev, c["log"], ev, rels, c["claimed"].words(), mine,
                                    [acc[ev].at(k << ev) for k in range(4)], lut_ptrs, accumulate=not fresh)
    comp, comp_log = None, 0
    for ev in sorted(acc):  # DomainEvaluationAccumulator::finalize: lift the smaller ones, interpolate
        vals = acc[ev]
        if comp is not None:
            lifted = be.alloc(4 << ev)
            be.evaluate(ColumnBatch(comp, 4, comp_log), ColumnBatch(lifted, 4, ev))
            be.accumulate([vals.at(k << ev) for k in range(4)], [lifted.at(k << ev) for k in range(4)], 1 << ev)
            keep.append(lifted)
        be.interpolate(ColumnBatch(vals, 4, ev))
        comp, comp_log = vals, ev
    _commit(be, trees[3], [(comp, 4, comp_log)], blowup, ch)

    # ---- OODS point, mask points, sampled values
    t = ch.draw_secure_felt()
    t2 = t * t
    inv = (t2 + ONE).inv()
    oods = ((ONE - t2) * inv, (t + t) * inv)
    points = [[[] for _ in tr.polys] for tr in trees]
    for slot in order:
        c = comps[slot]
        for i in c["pre_idx"]:
            points[0][i] = [oods]
        for k in range(c["n_main"]):
            points[1][c["main_loc"] + k] = [oods]
        sp = index_to_point(-subgroup_gen(c["log"]))
        prev = qpt_add(oods, (QM31(sp[0]), QM31(sp[1])))
        n_ic = 4 * c["n_fracs"]
        for k in range(n_ic):
            points[2][c["inter_loc"] + k] = [prev, oods] if k >= n_ic - 4 else [oods]
    points[3] = [[oods] for _ in range(4)]
    sampled = [[[None] * len(pts) for pts in tr] for tr in points]
    jobs = {}
    for ti, tr in enumerate(points):
        for ci, pts in enumerate(tr):
            for si, pt in enumerate(pts):
                ptr, log = trees[ti].polys[ci]
                jobs.setdefault((log, pt[0].c + pt[1].c), []).append((ptr, ti, ci, si))
    for (log, key), items in jobs.items():
        res = be.eval_at_point([it[0] for it in items], log, key)
        for (_, ti, ci, si), v in zip(items, res):
            sampled[ti][ci][si] = QM31(*[int(x) for x in v])
    ch.mix_felts([v for tr in sampled for col in tr for v in col])
    rc = ch.draw_secure_felt()

    # ---- DEEP quotients, one QM31 column per committed size
    flat = [(ptr, log, points[ti][ci], sampled[ti][ci]) for ti in range(4) for ci, (ptr, log) in enumerate(trees[ti].ldes)]
    quotients = []
    for log in sorted({f[1] for f in flat}, reverse=True):
        grp = [f for f in flat if f[1] == log]
        batches = {}
        for ci, (_, _, pts, vals) in enumerate(grp):
            for pt, v in zip(pts, vals):
                batches.setdefault(pt[0].c + pt[1].c, []).append((ci, v.words()))
        out = be.alloc(4 << log)
        be.accumulate_quotients(log, [f[0] for f in grp], list(batches.items()), rc.words(), [out.at(k << log) for k in range(4)])
        quotients.append((log, out))

    # ---- FRI commit
    fri_first = _Tree()
    for log, buf in quotients:
        fri_first.ldes += [(buf.at(k << log), log) for k in range(4)]
    _merkle_commit(be, fri_first)
    ch.mix_root(fri_first.root)
    alpha = ch.draw_secure_felt()
    line_log = quotients[0][0] - 1
    last_log = cfg.log_last_layer_degree_bound + blowup
    cur = be.alloc(4 << line_log)
    be.lib.lb_memset_zero(be.ctx, cur.ptr, 4 << line_log)
    inner, qi = [], 0
    while line_log > last_log:
        coords = [cur.at(k << line_log) for k in range(4)]
        while qi < len(quotients) and quotients[qi][0] - 1 == line_log:
            qlog, qbuf = quotients[qi]
            be.fold_circle_into_line(coords, [qbuf.at(k << qlog) for k in range(4)], qlog, alpha.words())
            qi += 1
        layer = _Tree()
        layer.ldes = [(p, line_log) for p in coords]
        layer.keep.append(cur)
        _merkle_commit(be, layer)
        ch.mix_root(layer.root)
        alpha = ch.draw_secure_felt()
        inner.append((line_log, coords, layer))
        nxt = be.alloc(4 << (line_log - 1))
        be.fold_line([nxt.at(k << (line_log - 1)) for k in range(4)], coords, line_log, alpha.words())
        cur, line_log = nxt, line_log - 1
    # last layer: LineEvaluation::interpolate on the host (2^(bound + blowup) values)
    n = 1 << line_log
    hv = be.download(cur, 4 * n).reshape(4, n)
    vals = [QM31(*[int(hv[k][bit_reverse(i, line_log)]) for k in range(4)]) for i in range(n)]
    d_init, d_step, size = subgroup_gen(line_log + 2), subgroup_gen(line_log), n
    while size > 1:
        for start in range(0, n, size):
            for i in range(size // 2):
                xinv = m_inv(index_to_point(d_init + d_step * i)[0])
                l, r = vals[start + i], vals[start + size // 2 + i]
                vals[start + i], vals[start + size // 2 + i] = l + r, (l - r) * xinv
        d_init, d_step, size = d_init * 2, d_step * 2, size // 2
    inv_n = m_inv(n % P)
    coeffs = [vals[bit_reverse(i, line_log)] * inv_n for i in range(n)]
    bound = 1 << cfg.log_last_layer_degree_bound
    if any(not c.is_zero() for c in coeffs[bound:]):
        raise ProvingError("fri: invalid last-layer degree (ConstraintsNotSatisfied)")
    last_poly = coeffs[:bound]
    ch.mix_felts(last_poly)

    # ---- proof of work, queries, decommitment
    nonce = be.grind(ch.digest, cfg.pow_bits, {"legacy": 0, "v2": 1}[channel_variant])
    ch.mix_u64(nonce)
    max_lde = quotients[0][0]
    qs, cnt = set(), 0
    while cnt < cfg.n_queries:
        rb = ch.draw_random_bytes()
        for k in range(8):
            if cnt < cfg.n_queries:
                qs.add(int.from_bytes(rb[4 * k: 4 * k + 4], "little") & ((1 << max_lde) - 1))
                cnt += 1
    queries = sorted(qs)
    qpos = {log: _fold_queries(queries, max_lde - log) for log, _ in quotients}
    pos_by_size, first_witness = {}, []
    for log, buf in quotients:
        positions, w = _fri_positions_and_witness(be, [buf.at(k << log) for k in range(4)], _fold_queries(queries, max_lde - log))
        pos_by_size[log] = positions
        first_witness += w
    _, dec = _decommit(be, fri_first, pos_by_size)
    first_layer = wire.FriLayer(first_witness, dec, fri_first.root)
    inner_layers, lq = [], _fold_queries(queries, 1)
    for log, coords, layer in inner:
        positions, w = _fri_positions_and_witness(be, coords, lq)
        _, dec = _decommit(be, layer, {log: positions})
        inner_layers.append(wire.FriLayer(w, dec, layer.root))
        lq = _fold_queries(lq, 1)
    queried, decommitments = [], []
    for tr in trees:
        qv, dec = _decommit(be, tr, qpos)
        queried.append(qv)
        decommitments.append(dec)

    # ---- OODS check (prove() returns ConstraintsNotSatisfied otherwise): left to the verifier in this twin - the composition
    # polynomial it would be compared with comes from the same lb_constraint_quotients calls lb_prove makes
    iclaim = [None] * n_slots
    for slot in order:
        iclaim[slot] = comps[slot]["claimed"].c
    p = wire.Proof(claim, iclaim, cfg.pow_bits, blowup, cfg.log_last_layer_degree_bound, cfg.n_queries,
                   [t.root for t in trees], [[[v.c for v in col] for col in tr] for tr in sampled], decommitments, queried, nonce,
                   first_layer, inner_layers, [c.c for c in last_poly], max(len(last_poly).bit_length() - 1, 0))
    return wire.to_bincode(p)


class _Raw:
    """A device pointer dressed as a DeviceBuffer for ColumnBatch."""

    def __init__(self, ptr: int):
        self.ptr = ptr

    def at(self, offset_u32: int) -> int:
        return self.ptr + 4 * int(offset_u32)
