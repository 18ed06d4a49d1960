// CPU prover (oracle; TEST INFRASTRUCTURE ONLY).
//
// LuminAIR's 17 component AIRs over a restatement of stwo-constraint-framework @0790eba (EvalAtRow with its three
// evaluators - InfoEvaluator, PointEvaluator, SimdDomainEvaluator - finalize_logup, relation!).  Written from the
// in-tree sources, one function per /root/reference/crates/air/src/components/<name>/component.rs::evaluate:
//   add :38-116   mul :40-128   recip :39-107   sin :51-123   sum_reduce :37-110   max_reduce :37-121   sqrt :38-107
//   rem :38-124   exp2 :46-118  log2 :46-117    less_than :49-184   inputs/components.rs:37-85   contiguous :37-101
//   lookups/{sin,exp2,log2}/component.rs:41-59   lookups/range_check/component.rs:44-60
// numerair's EvalFixedPoint helpers (un-vendored, rev 11d1d26) are restated as "dividend - (quotient * divisor + rem)";
// eval_fixed_add / eval_fixed_mul are pinned by the OODS check of the reference's committed proof.
#pragma once
#include <initializer_list>
#include <stdexcept>
#include <utility>
#include <vector>

#include "circle.hpp"

namespace cpu {

enum Rel { REL_NODE = 0, REL_SIN, REL_EXP2, REL_LOG2, REL_RANGE_CHECK, REL_COUNT };
// = field index in LuminairClaim (/root/reference/crates/air/src/lib.rs:30-48)
enum Kind {
    K_ADD = 0, K_MUL, K_RECIP, K_SIN, K_SIN_LOOKUP, K_SUM_REDUCE, K_MAX_REDUCE, K_SQRT, K_REM, K_EXP2, K_EXP2_LOOKUP,
    K_LOG2, K_LOG2_LOOKUP, K_LESS_THAN, K_RANGE_CHECK_LOOKUP, K_INPUTS, K_CONTIGUOUS, K_COUNT
};
constexpr uint32_t FP_SCALE = 1u << 12;  // DEFAULT_FP_SCALE, crates/air/src/lib.rs:23

// relation!(Name, N): combine(v) = sum_i alpha^i v_i - z
struct Relation {
    QM z, alpha, pw[2];
    void set(QM z_, QM a_) {
        z = z_;
        alpha = a_;
        pw[0] = qm(1);
        pw[1] = a_;
    }
};

struct CompInfo {
    int n_main, n_frac;  // main columns, LogUp fractions (= extension interaction columns)
    int lut;             // Rel of the lookup table the component consumes / tabulates, -1 if none
    int n_lut_cols;      // table components: preprocessed columns read (2; range check 1), else 0
    std::vector<std::pair<int, uint32_t>> padding;  // non-zero entries of the padding row (<name>/table.rs padding())
};
static inline const CompInfo& comp_info(int kind) {
    static const CompInfo T[K_COUNT] = {
        /* add */ {15, 3, -1, 0, {{4, 1}}},
        /* mul */ {16, 3, -1, 0, {{4, 1}}},
        /* recip */ {13, 2, -1, 0, {{3, 1}}},
        /* sin */ {12, 3, REL_SIN, 0, {{3, 1}}},
        /* sin_lookup */ {1, 1, REL_SIN, 2, {}},
        /* sum_reduce */ {14, 2, -1, 0, {{3, 1}}},
        /* max_reduce */ {15, 2, -1, 0, {{3, 1}}},
        /* sqrt */ {13, 2, -1, 0, {{3, 1}}},
        /* rem */ {16, 3, -1, 0, {{4, 1}}},
        /* exp2 */ {12, 3, REL_EXP2, 0, {{3, 1}}},
        /* exp2_lookup */ {1, 1, REL_EXP2, 2, {}},
        /* log2 */ {12, 3, REL_LOG2, 0, {{3, 1}}},
        /* log2_lookup */ {1, 1, REL_LOG2, 2, {}},
        /* less_than: 0 < 1 -> out = 1.0, diff = 1, limb0 = 1 (less_than/table.rs padding()) */
        {22, 7, REL_RANGE_CHECK, 0, {{4, 1}, {10, 1}, {11, FP_SCALE}, {12, 1}, {14, 1}}},
        /* range_check_lookup */ {1, 1, REL_RANGE_CHECK, 1, {}},
        /* inputs */ {7, 1, -1, 0, {{2, 1}}},
        /* contiguous */ {11, 2, -1, 0, {{3, 1}}},
    };
    return T[kind];
}

struct CompCtx {
    int kind;
    int log_size;
    const Relation* rels;  // REL_COUNT
    int legacy_mul_extra;  // the Mul AIR of the revision behind ui/demo/public/proof had one more (vanishing) constraint
};

// ---- value types of the three evaluators --------------------------------------------------------------------------------
struct FV { V v; };
struct EV { VQ q; };
static inline FV operator+(FV a, FV b) { return {a.v + b.v}; }
static inline FV operator-(FV a, FV b) { return {a.v - b.v}; }
static inline FV operator*(FV a, FV b) { return {a.v * b.v}; }
static inline EV operator+(EV a, EV b) { return {a.q + b.q}; }
static inline EV operator-(EV a, EV b) { return {a.q - b.q}; }
static inline EV operator*(EV a, EV b) { return {a.q * b.q}; }
static inline EV operator-(EV a, FV b) {
    a.q.c[0] = a.q.c[0] - b.v;
    return a;
}
struct QS { QM q; };
static inline QS operator+(QS a, QS b) { return {a.q + b.q}; }
static inline QS operator-(QS a, QS b) { return {a.q - b.q}; }
static inline QS operator*(QS a, QS b) { return {a.q * b.q}; }
struct Nil {};
static inline Nil operator+(Nil, Nil) { return {}; }
static inline Nil operator-(Nil, Nil) { return {}; }
static inline Nil operator*(Nil, Nil) { return {}; }

// ---- LogUp part shared by the evaluators (constraint-framework logup.rs: one extension column per fraction, the
// last one with the [-1, 0] mask and the cumulative-sum shift) -------------------------------------------------------------
template <class E, class F, class EF>
struct LogupMixin {
    std::vector<std::pair<F, EF>> fracs;
    E& self() { return *static_cast<E*>(this); }
    void add_to_relation(const Relation& r, F mult, std::initializer_list<F> values) {
        int i = 0;
        EF acc = self().econst(qm(0));
        for (const F& v : values) acc = acc + self().mulc(r.pw[i++], v);
        fracs.push_back({mult, acc - self().econst(r.z)});
    }
    void finalize_logup() {
        bool have_prev = false;
        EF prev_col = self().econst(qm(0));
        for (size_t k = 0; k + 1 < fracs.size(); k++) {
            EF cur = self().next_ext_mask();
            EF diff = have_prev ? cur - prev_col : cur;
            prev_col = cur;
            have_prev = true;
            self().add_constraint_ext(self().sub_f(diff * fracs[k].second, fracs[k].first));
        }
        EF prev_row = self().econst(qm(0)), cur = prev_row;
        self().next_ext_mask_prev_cur(prev_row, cur);
        EF diff = cur - prev_row;
        if (have_prev) diff = diff - prev_col;
        EF fixed = diff + self().econst(self().cumsum_shift());
        self().add_constraint_ext(self().sub_f(fixed * fracs.back().second, fracs.back().first));
        fracs.clear();
    }
};

struct InfoEval : LogupMixin<InfoEval, Nil, Nil> {
    using F = Nil;
    using EF = Nil;
    int n_constraints = 0, n_main = 0, n_ext = 0, n_pre = 0;
    F next() { n_main++; return {}; }
    F pre(int) { n_pre++; return {}; }
    F c(uint32_t) { return {}; }
    EF econst(QM) { return {}; }
    EF mulc(QM, F) { return {}; }
    EF sub_f(EF, F) { return {}; }
    EF next_ext_mask() { n_ext++; return {}; }
    void next_ext_mask_prev_cur(EF&, EF&) { n_ext++; }
    QM cumsum_shift() { return qm(0); }
    void add_constraint(F) { n_constraints++; }
    void add_constraint_ext(EF) { n_constraints++; }
};

// PointEvaluator: mask values are the OODS samples; accumulates acc = acc * r + denom_inverse * constraint
struct PointEval : LogupMixin<PointEval, QS, QS> {
    using F = QS;
    using EF = QS;
    const std::vector<QM>* main;   // per main column of this component: samples
    const std::vector<QM>* inter;  // per interaction base column
    const std::vector<QM>* const* prep;  // per preprocessed column the component reads
    int mi = 0, ii = 0;
    QM denom_inverse, shift, r;
    QM* acc;
    F next() { return {main[mi++][0]}; }
    F pre(int k) { return {(*prep[k])[0]}; }
    F c(uint32_t x) { return {qm(x)}; }
    EF econst(QM q) { return {q}; }
    EF mulc(QM p, F v) { return {p * v.q}; }
    EF sub_f(EF a, F b) { return {a.q - b.q}; }
    QM ext(int sample) {
        QM e[4] = {inter[ii][sample], inter[ii + 1][sample], inter[ii + 2][sample], inter[ii + 3][sample]};
        return from_partial_evals(e);
    }
    EF next_ext_mask() {
        QM v = ext(0);
        ii += 4;
        return {v};
    }
    void next_ext_mask_prev_cur(EF& prev, EF& cur) {
        prev.q = ext(0);
        cur.q = ext(1);
        ii += 4;
    }
    QM cumsum_shift() { return shift; }
    void add_constraint(F x) { *acc = *acc * r + denom_inverse * x.q; }
    void add_constraint_ext(EF x) { add_constraint(x); }
};

// SimdDomainEvaluator: W rows of the evaluation domain at a time
struct DomainEval : LogupMixin<DomainEval, FV, EV> {
    using F = FV;
    using EF = EV;
    const uint32_t* const* main;
    const uint32_t* const* inter;
    const uint32_t* const* prep;
    const uint32_t* prev_idx;  // storage row of the [-1] mask for every row of the evaluation domain
    size_t row = 0;
    int mi = 0, ii = 0, ci = 0;
    const QM* pows;  // first constraint first
    QM shift;
    VQ row_res;
    F next() { return {vload(main[mi++] + row)}; }
    F pre(int k) { return {vload(prep[k] + row)}; }
    F c(uint32_t x) { return {vset1(x)}; }
    EF econst(QM q) { return {vq_set1(q)}; }
    EF mulc(QM p, F v) { return {VQ{{vset1(p.c[0]) * v.v, vset1(p.c[1]) * v.v, vset1(p.c[2]) * v.v, vset1(p.c[3]) * v.v}}}; }
    EF sub_f(EF a, F b) { return a - b; }
    EF next_ext_mask() {
        EV e{VQ{{vload(inter[ii] + row), vload(inter[ii + 1] + row), vload(inter[ii + 2] + row), vload(inter[ii + 3] + row)}}};
        ii += 4;
        return e;
    }
    void next_ext_mask_prev_cur(EF& prev, EF& cur) {
        for (int k = 0; k < 4; k++) {
            uint32_t t[W];
            for (int l = 0; l < W; l++) t[l] = inter[ii + k][prev_idx[row + l]];
            prev.q.c[k] = vload(t);
            cur.q.c[k] = vload(inter[ii + k] + row);
        }
        ii += 4;
    }
    QM cumsum_shift() { return shift; }
    void add_constraint(F x) {
        QM p = pows[ci++];
        for (int k = 0; k < 4; k++) row_res.c[k] = row_res.c[k] + vset1(p.c[k]) * x.v;
    }
    void add_constraint_ext(EF x) { row_res = row_res + vq_set1(pows[ci++]) * x.q; }
};

// ---- the AIRs ----------------------------------------------------------------------------------------------------------------
template <class E>
static void eval_binary(E& ev, const CompCtx& cx) {  // add / mul / rem / less_than share the 9-column head
    using F = typename E::F;
    const Relation& node = cx.rels[REL_NODE];
    F node_id = ev.next(), lhs_id = ev.next(), rhs_id = ev.next(), idx = ev.next(), is_last_idx = ev.next();
    F next_node_id = ev.next(), next_lhs_id = ev.next(), next_rhs_id = ev.next(), next_idx = ev.next();
    F lhs_val = ev.next(), rhs_val = ev.next();
    F one = ev.c(1);
    auto transitions = [&] {
        F not_last = one - is_last_idx;
        ev.add_constraint(not_last * (next_node_id - node_id));
        ev.add_constraint(not_last * (next_lhs_id - lhs_id));
        ev.add_constraint(not_last * (next_rhs_id - rhs_id));
        ev.add_constraint(not_last * (next_idx - idx - one));
    };
    if (cx.kind == K_ADD) {
        F out_val = ev.next(), lhs_mult = ev.next(), rhs_mult = ev.next(), out_mult = ev.next();
        ev.add_constraint(is_last_idx * (is_last_idx - one));
        ev.add_constraint(out_val - (lhs_val + rhs_val));  // eval_fixed_add
        transitions();
        ev.add_to_relation(node, lhs_mult, {lhs_val, lhs_id});
        ev.add_to_relation(node, rhs_mult, {rhs_val, rhs_id});
        ev.add_to_relation(node, out_mult, {out_val, node_id});
    } else if (cx.kind == K_MUL) {
        F out_val = ev.next(), rem_val = ev.next(), lhs_mult = ev.next(), rhs_mult = ev.next(), out_mult = ev.next();
        ev.add_constraint(is_last_idx * (is_last_idx - one));
        ev.add_constraint(lhs_val * rhs_val - (out_val * ev.c(FP_SCALE) + rem_val));  // eval_fixed_mul
        for (int k = 0; k < cx.legacy_mul_extra; k++) ev.add_constraint(rem_val * ev.c(0));
        transitions();
        ev.add_to_relation(node, lhs_mult, {lhs_val, lhs_id});
        ev.add_to_relation(node, rhs_mult, {rhs_val, rhs_id});
        ev.add_to_relation(node, out_mult, {out_val, node_id});
    } else if (cx.kind == K_REM) {
        F rem_val = ev.next(), quotient = ev.next(), lhs_mult = ev.next(), rhs_mult = ev.next(), out_mult = ev.next();
        ev.add_constraint(is_last_idx * (is_last_idx - one));
        ev.add_constraint(lhs_val - (quotient * rhs_val + rem_val));  // eval_fixed_rem
        transitions();
        ev.add_to_relation(node, lhs_mult, {lhs_val, lhs_id});
        ev.add_to_relation(node, rhs_mult, {rhs_val, rhs_id});
        ev.add_to_relation(node, out_mult, {rem_val, node_id});
    } else {  // K_LESS_THAN
        const Relation& rc = cx.rels[REL_RANGE_CHECK];
        F out_val = ev.next(), diff_val = ev.next(), borrow = ev.next();
        F limb0 = ev.next(), limb1 = ev.next(), limb2 = ev.next(), limb3 = ev.next();
        F lhs_mult = ev.next(), rhs_mult = ev.next(), out_mult = ev.next(), diff_mult = ev.next();
        ev.add_constraint(is_last_idx * (is_last_idx - one));
        ev.add_constraint(borrow * (borrow - one));
        ev.add_constraint(out_val - ((one - borrow) * ev.c(FP_SCALE)));
        // the reference passes TWO_POW_31_MINUS_1 (crates/air/src/lib.rs:26) as "2^k": 0 in M31
        ev.add_constraint(lhs_val + diff_val - rhs_val - (borrow * ev.c(0)));
        F recomposed = limb3 * ev.c(1u << 24) + limb2 * ev.c(1u << 16) + limb1 * ev.c(1u << 8) + limb0;
        ev.add_constraint(diff_val - recomposed);
        transitions();
        ev.add_to_relation(node, lhs_mult, {lhs_val, lhs_id});
        ev.add_to_relation(node, rhs_mult, {rhs_val, rhs_id});
        ev.add_to_relation(node, out_mult, {out_val, node_id});
        ev.add_to_relation(rc, diff_mult, {limb0});
        ev.add_to_relation(rc, diff_mult, {limb1});
        ev.add_to_relation(rc, diff_mult, {limb2});
        ev.add_to_relation(rc, diff_mult, {limb3});
    }
    ev.finalize_logup();
}

template <class E>
static void eval_unary(E& ev, const CompCtx& cx) {  // the 7-column head: node_id, input_id, idx, is_last_idx, next_*
    using F = typename E::F;
    const Relation& node = cx.rels[REL_NODE];
    F node_id = ev.next(), input_id = ev.next(), idx = ev.next(), is_last_idx = ev.next();
    F next_node_id = ev.next(), next_input_id = ev.next(), next_idx = ev.next();
    F one = ev.c(1);
    auto transitions = [&] {
        F not_last = one - is_last_idx;
        ev.add_constraint(not_last * (next_node_id - node_id));
        ev.add_constraint(not_last * (next_input_id - input_id));
        ev.add_constraint(not_last * (next_idx - idx - one));
    };
    switch (cx.kind) {
        case K_SUM_REDUCE: {
            F input_val = ev.next(), out_val = ev.next(), acc_val = ev.next(), next_acc_val = ev.next();
            F is_last_step = ev.next(), input_mult = ev.next(), out_mult = ev.next();
            ev.add_constraint(is_last_idx * (is_last_idx - one));
            ev.add_constraint(is_last_step * (is_last_step - one));
            ev.add_constraint(next_acc_val - (acc_val + input_val));
            ev.add_constraint((out_val - next_acc_val) * is_last_step);
            transitions();
            ev.add_to_relation(node, input_mult, {input_val, input_id});
            ev.add_to_relation(node, out_mult, {out_val, node_id});
            break;
        }
        case K_MAX_REDUCE: {
            F input_val = ev.next(), out_val = ev.next(), max_val = ev.next(), next_max_val = ev.next();
            F is_last_step = ev.next(), is_max = ev.next(), input_mult = ev.next(), out_mult = ev.next();
            ev.add_constraint(is_last_idx * (is_last_idx - one));
            ev.add_constraint(is_last_step * (is_last_step - one));
            ev.add_constraint(is_max * (is_max - one));
            ev.add_constraint(is_max * (next_max_val - input_val));
            ev.add_constraint((one - is_max) * (next_max_val - max_val));
            ev.add_constraint((out_val - next_max_val) * is_last_step);
            transitions();
            ev.add_to_relation(node, input_mult, {input_val, input_id});
            ev.add_to_relation(node, out_mult, {out_val, node_id});
            break;
        }
        case K_CONTIGUOUS: {
            F inp = ev.next(), out = ev.next(), input_mult = ev.next(), out_mult = ev.next();
            ev.add_constraint(is_last_idx * (is_last_idx - one));
            transitions();
            ev.add_to_relation(node, input_mult, {inp, input_id});
            ev.add_to_relation(node, out_mult, {out, node_id});
            break;
        }
        case K_RECIP:
        case K_SQRT: {
            F input_val = ev.next(), out_val = ev.next(), rem_val = ev.next(), scale = ev.next();
            F input_mult = ev.next(), out_mult = ev.next();
            ev.add_constraint(is_last_idx * (is_last_idx - one));
            if (cx.kind == K_RECIP)
                ev.add_constraint(scale * scale - (input_val * out_val + rem_val));  // eval_fixed_recip
            else
                ev.add_constraint(input_val * scale - (out_val * out_val + rem_val));  // eval_fixed_sqrt
            transitions();
            ev.add_to_relation(node, input_mult, {input_val, input_id});
            ev.add_to_relation(node, out_mult, {out_val, node_id});
            break;
        }
        default: {  // K_SIN, K_EXP2, K_LOG2: third relation use = (input, output) against the function's table
            F input_val = ev.next(), out_val = ev.next(), input_mult = ev.next(), out_mult = ev.next(), lookup_mult = ev.next();
            ev.add_constraint(is_last_idx * (is_last_idx - one));
            transitions();
            ev.add_to_relation(node, input_mult, {input_val, input_id});
            ev.add_to_relation(node, out_mult, {out_val, node_id});
            ev.add_to_relation(cx.rels[comp_info(cx.kind).lut], lookup_mult, {input_val, out_val});
        }
    }
    ev.finalize_logup();
}

template <class E>
static void evaluate_component(E& ev, const CompCtx& cx) {
    using F = typename E::F;
    switch (cx.kind) {
        case K_ADD: case K_MUL: case K_REM: case K_LESS_THAN:
            eval_binary(ev, cx);
            break;
        case K_INPUTS: {
            F node_id = ev.next(), idx = ev.next(), is_last_idx = ev.next(), next_node_id = ev.next(), next_idx = ev.next();
            F val = ev.next(), multiplicity = ev.next();
            F one = ev.c(1);
            ev.add_constraint(is_last_idx * (is_last_idx - one));
            F not_last = one - is_last_idx;
            ev.add_constraint(not_last * (next_node_id - node_id));
            ev.add_constraint(not_last * (next_idx - idx - one));
            ev.add_to_relation(cx.rels[REL_NODE], multiplicity, {val, node_id});
            ev.finalize_logup();
            break;
        }
        case K_SIN_LOOKUP: case K_EXP2_LOOKUP: case K_LOG2_LOOKUP: {
            F l0 = ev.pre(0), l1 = ev.pre(1);
            F multiplicity = ev.next();
            ev.add_to_relation(cx.rels[comp_info(cx.kind).lut], ev.c(0) - multiplicity, {l0, l1});
            ev.finalize_logup();
            break;
        }
        case K_RANGE_CHECK_LOOKUP: {
            F l0 = ev.pre(0);
            F multiplicity = ev.next();
            ev.add_to_relation(cx.rels[REL_RANGE_CHECK], ev.c(0) - multiplicity, {l0});
            ev.finalize_logup();
            break;
        }
        default:
            eval_unary(ev, cx);
    }
}

// LogUp fractions of the interaction-trace writers (<name>/witness.rs write_interaction_trace, e.g.
// add/witness.rs:126-167, less_than/witness.rs:144-226, lookups/exp2/witness.rs:117-144): multiplicity column, value
// columns (main-trace column indices; table components: preprocessed columns with the multiplicity negated), relation
struct LookupTerm {
    int mult, nvals, val[2], rel;
    bool table;
};
static inline std::vector<LookupTerm> lookup_terms(int kind) {
    auto N = REL_NODE;
    switch (kind) {
        case K_ADD: return {{12, 2, {9, 1}, N, false}, {13, 2, {10, 2}, N, false}, {14, 2, {11, 0}, N, false}};
        case K_MUL:
        case K_REM: return {{13, 2, {9, 1}, N, false}, {14, 2, {10, 2}, N, false}, {15, 2, {11, 0}, N, false}};
        case K_INPUTS: return {{6, 2, {5, 0}, N, false}};
        case K_SUM_REDUCE: return {{12, 2, {7, 1}, N, false}, {13, 2, {8, 0}, N, false}};
        case K_MAX_REDUCE: return {{13, 2, {7, 1}, N, false}, {14, 2, {8, 0}, N, false}};
        case K_CONTIGUOUS: return {{9, 2, {7, 1}, N, false}, {10, 2, {8, 0}, N, false}};
        case K_RECIP:
        case K_SQRT: return {{11, 2, {7, 1}, N, false}, {12, 2, {8, 0}, N, false}};
        case K_SIN:
        case K_EXP2:
        case K_LOG2: return {{9, 2, {7, 1}, N, false}, {10, 2, {8, 0}, N, false}, {11, 2, {7, 8}, comp_info(kind).lut, false}};
        case K_LESS_THAN:
            return {{18, 2, {9, 1}, N, false}, {19, 2, {10, 2}, N, false}, {20, 2, {11, 0}, N, false},
                    {21, 1, {14, 0}, REL_RANGE_CHECK, false}, {21, 1, {15, 0}, REL_RANGE_CHECK, false},
                    {21, 1, {16, 0}, REL_RANGE_CHECK, false}, {21, 1, {17, 0}, REL_RANGE_CHECK, false}};
        case K_SIN_LOOKUP:
        case K_EXP2_LOOKUP:
        case K_LOG2_LOOKUP: return {{0, 2, {0, 1}, comp_info(kind).lut, true}};
        case K_RANGE_CHECK_LOOKUP: return {{0, 1, {0, 0}, REL_RANGE_CHECK, true}};
    }
    throw std::runtime_error("unknown component");
}

}  // namespace cpu
