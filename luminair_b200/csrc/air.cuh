// LuminAIR component AIRs written once against an abstract row evaluator, the way the reference
// writes `FrameworkEval::evaluate<E: EvalAtRow>`:
//   Add     /root/reference/crates/air/src/components/add/component.rs:38-116
//   Mul     /root/reference/crates/air/src/components/mul/component.rs:40-128
//   Inputs  /root/reference/crates/air/src/components/inputs/components.rs:37-85
// plus the LogUp bookkeeping of stwo-constraint-framework (`add_to_relation`, `finalize_logup`,
// un-vendored, rev 0790eba) and numerair's `eval_fixed_add` / `eval_fixed_mul` (rev 11d1d26).
//
// Three evaluators instantiate these templates:
//   InfoEval    (host)   counts constraints / mask offsets                       -> prover.cu
//   PointEval   (host)   F = EF = QM31, evaluates at the OODS point              -> prover.cu
//   DomainEval  (device) F = M31, EF = QM31, one thread per evaluation-domain row -> air_kernels.cu
#pragma once
#include "m31.cuh"

namespace lb {

// ---- field wrappers with operators ---------------------------------------------------------
struct FM {
    uint32_t v;
};
__host__ __device__ __forceinline__ FM operator+(FM a, FM b) { return {m_add(a.v, b.v)}; }
__host__ __device__ __forceinline__ FM operator-(FM a, FM b) { return {m_sub(a.v, b.v)}; }
__host__ __device__ __forceinline__ FM operator*(FM a, FM b) { return {m_mul(a.v, b.v)}; }

struct FQ {
    QM31 v;
};
__host__ __device__ __forceinline__ FQ operator+(FQ a, FQ b) { return {q_add(a.v, b.v)}; }
__host__ __device__ __forceinline__ FQ operator-(FQ a, FQ b) { return {q_sub(a.v, b.v)}; }
__host__ __device__ __forceinline__ FQ operator*(FQ a, FQ b) { return {q_mul(a.v, b.v)}; }
// mixed
__host__ __device__ __forceinline__ FQ operator*(FQ a, FM b) { return {q_mul_m(a.v, b.v)}; }
__host__ __device__ __forceinline__ FQ operator+(FQ a, FM b) {
    QM31 r = a.v;
    r.a.a = m_add(r.a.a, b.v);
    return {r};
}
__host__ __device__ __forceinline__ FQ operator-(FQ a, FM b) {
    QM31 r = a.v;
    r.a.a = m_sub(r.a.a, b.v);
    return {r};
}
__host__ __device__ __forceinline__ FQ to_ef(FM a) { return {q_from_m(a.v)}; }
__host__ __device__ __forceinline__ FQ to_ef(FQ a) { return a; }

// relation!(NodeElements, 2) (crates/air/src/components/mod.rs:218): combine(v) = sum alpha^i v_i - z
struct Relation2 {
    QM31 z;
    QM31 alpha;  // alpha^1 (alpha^0 = 1)
};

constexpr uint32_t FP_SCALE = 1u << 12;  // DEFAULT_FP_SCALE, crates/air/src/lib.rs:23
constexpr int MAX_FRACS = 4;

// LogUp bookkeeping shared by the evaluators (CRTP: E provides F, EF, the mask readers,
// add_constraint_ef and the `cumsum_shift` member).
template <class E, class F, class EF>
struct LogupMixin {
    F num[MAX_FRACS];
    EF den[MAX_FRACS];
    int n_fracs = 0;

#pragma nv_exec_check_disable
    __host__ __device__ __forceinline__ void add_to_relation(const Relation2& rel, F multiplicity, F v0, F v1) {
        EF a = EF{rel.alpha} * v1;
        EF d = (a + v0) - EF{rel.z};
        num[n_fracs] = multiplicity;
        den[n_fracs] = d;
        ++n_fracs;
    }

    // one interaction (QM31) column per fraction; the last one carries the [-1, 0] mask and the
    // cumulative-sum shift.
#pragma nv_exec_check_disable
    __host__ __device__ __forceinline__ void finalize_logup() {
        E& self = *static_cast<E*>(this);
        EF prev_col{};
        bool have_prev = false;
#pragma unroll
        for (int k = 0; k < MAX_FRACS; ++k) {
            if (k >= n_fracs - 1) break;
            EF cur = self.next_ext_mask_cur();
            EF diff = have_prev ? (cur - prev_col) : cur;
            prev_col = cur;
            have_prev = true;
            self.add_constraint_ef(diff * den[k] - num[k]);
        }
        EF prev_row, cur;
        self.next_ext_mask_prev_cur(prev_row, cur);
        EF diff = cur - prev_row;
        if (have_prev) diff = diff - prev_col;
        EF fixed = diff + EF{self.cumsum_shift};
        self.add_constraint_ef(fixed * den[n_fracs - 1] - num[n_fracs - 1]);
        n_fracs = 0;
    }
};

// ---- components ------------------------------------------------------------------------------
// COMP_MUL_ARTIFACT: the Mul AIR of the LuminAIR revision that produced the reference's committed proof
// (ui/demo/public/proof): one more constraint slot after eval_fixed_mul, identically zero.  It exists
// only so the known-answer test can replay that proof byte-for-byte.
enum ComponentKind {
    COMP_ADD = 0,
    COMP_MUL = 1,
    COMP_INPUTS = 2,
    COMP_MUL_ARTIFACT = 3,
    COMP_SUM_REDUCE = 4,  // components/sum_reduce/component.rs:37-110
    COMP_MAX_REDUCE = 5,  // components/max_reduce/component.rs:37-121
    COMP_CONTIGUOUS = 6,  // components/contiguous/component.rs:37-101
    COMP_KIND_COUNT = 7
};

struct ComponentShape {
    int n_main;         // main-trace columns (add/witness.rs:24 etc.)
    int n_fracs;        // LogUp relation uses -> n_fracs QM31 interaction columns
    int n_constraints;  // counted by InfoEval at start-up, checked against this table
    int padding_one_col;  // index of `is_last_idx`, the only non-zero entry of the padding row
};
__host__ __device__ constexpr ComponentShape component_shape(int kind) {
    return kind == COMP_ADD            ? ComponentShape{15, 3, 9, 4}
           : kind == COMP_MUL          ? ComponentShape{16, 3, 9, 4}
           : kind == COMP_MUL_ARTIFACT ? ComponentShape{16, 3, 10, 4}
           : kind == COMP_SUM_REDUCE   ? ComponentShape{14, 2, 9, 3}
           : kind == COMP_MAX_REDUCE   ? ComponentShape{15, 2, 11, 3}
           : kind == COMP_CONTIGUOUS   ? ComponentShape{11, 2, 6, 3}
                                       : ComponentShape{7, 1, 4, 2};
}

#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_add(E& ev, const Relation2& node) {
    typedef typename E::F F;
    F node_id = ev.next_trace_mask();
    F lhs_id = ev.next_trace_mask();
    F rhs_id = ev.next_trace_mask();
    F idx = ev.next_trace_mask();
    F is_last_idx = ev.next_trace_mask();
    F next_node_id = ev.next_trace_mask();
    F next_lhs_id = ev.next_trace_mask();
    F next_rhs_id = ev.next_trace_mask();
    F next_idx = ev.next_trace_mask();
    F lhs_val = ev.next_trace_mask();
    F rhs_val = ev.next_trace_mask();
    F out_val = ev.next_trace_mask();
    F lhs_mult = ev.next_trace_mask();
    F rhs_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F one = ev.constant(1);

    ev.add_constraint(is_last_idx * (is_last_idx - one));
    ev.add_constraint(out_val - (lhs_val + rhs_val));  // eval_fixed_add
    F not_last = one - is_last_idx;
    ev.add_constraint(not_last * (next_node_id - node_id));
    ev.add_constraint(not_last * (next_lhs_id - lhs_id));
    ev.add_constraint(not_last * (next_rhs_id - rhs_id));
    ev.add_constraint(not_last * (next_idx - idx - one));
    ev.add_to_relation(node, lhs_mult, lhs_val, lhs_id);
    ev.add_to_relation(node, rhs_mult, rhs_val, rhs_id);
    ev.add_to_relation(node, out_mult, out_val, node_id);
    ev.finalize_logup();
}

#pragma nv_exec_check_disable
template <class E, bool ARTIFACT = false>
__host__ __device__ __forceinline__ void eval_mul(E& ev, const Relation2& node) {
    typedef typename E::F F;
    F node_id = ev.next_trace_mask();
    F lhs_id = ev.next_trace_mask();
    F rhs_id = ev.next_trace_mask();
    F idx = ev.next_trace_mask();
    F is_last_idx = ev.next_trace_mask();
    F next_node_id = ev.next_trace_mask();
    F next_lhs_id = ev.next_trace_mask();
    F next_rhs_id = ev.next_trace_mask();
    F next_idx = ev.next_trace_mask();
    F lhs_val = ev.next_trace_mask();
    F rhs_val = ev.next_trace_mask();
    F out_val = ev.next_trace_mask();
    F rem_val = ev.next_trace_mask();
    F lhs_mult = ev.next_trace_mask();
    F rhs_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    F scale = ev.constant(FP_SCALE);

    ev.add_constraint(is_last_idx * (is_last_idx - one));
    ev.add_constraint(lhs_val * rhs_val - (out_val * scale + rem_val));  // eval_fixed_mul
    if (ARTIFACT) ev.add_constraint(rem_val * ev.constant(0));
    F not_last = one - is_last_idx;
    ev.add_constraint(not_last * (next_node_id - node_id));
    ev.add_constraint(not_last * (next_lhs_id - lhs_id));
    ev.add_constraint(not_last * (next_rhs_id - rhs_id));
    ev.add_constraint(not_last * (next_idx - idx - one));
    ev.add_to_relation(node, lhs_mult, lhs_val, lhs_id);
    ev.add_to_relation(node, rhs_mult, rhs_val, rhs_id);
    ev.add_to_relation(node, out_mult, out_val, node_id);
    ev.finalize_logup();
}

#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_inputs(E& ev, const Relation2& node) {
    typedef typename E::F F;
    F node_id = ev.next_trace_mask();
    F idx = ev.next_trace_mask();
    F is_last_idx = ev.next_trace_mask();
    F next_node_id = ev.next_trace_mask();
    F next_idx = ev.next_trace_mask();
    F val = ev.next_trace_mask();
    F multiplicity = ev.next_trace_mask();
    F one = ev.constant(1);

    ev.add_constraint(is_last_idx * (is_last_idx - one));
    F not_last = one - is_last_idx;
    ev.add_constraint(not_last * (next_node_id - node_id));
    ev.add_constraint(not_last * (next_idx - idx - one));
    ev.add_to_relation(node, multiplicity, val, node_id);
    ev.finalize_logup();
}

#pragma nv_exec_check_disable
// head shared by SumReduce / MaxReduce / Contiguous: ids, index, next-row copies
template <class E>
struct ReduceHead {
    typename E::F node_id, input_id, idx, is_last_idx, next_node_id, next_input_id, next_idx;
};
#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void read_reduce_head(E& ev, ReduceHead<E>& h) {
    h.node_id = ev.next_trace_mask();
    h.input_id = ev.next_trace_mask();
    h.idx = ev.next_trace_mask();
    h.is_last_idx = ev.next_trace_mask();
    h.next_node_id = ev.next_trace_mask();
    h.next_input_id = ev.next_trace_mask();
    h.next_idx = ev.next_trace_mask();
}
#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void reduce_transitions(E& ev, const ReduceHead<E>& h) {
    typedef typename E::F F;
    F one = ev.constant(1);
    F not_last = one - h.is_last_idx;
    ev.add_constraint(not_last * (h.next_node_id - h.node_id));
    ev.add_constraint(not_last * (h.next_input_id - h.input_id));
    ev.add_constraint(not_last * (h.next_idx - h.idx - one));
}

#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_sum_reduce(E& ev, const Relation2& node) {
    typedef typename E::F F;
    ReduceHead<E> h;
    read_reduce_head(ev, h);
    F input_val = ev.next_trace_mask();
    F out_val = ev.next_trace_mask();
    F acc_val = ev.next_trace_mask();
    F next_acc_val = ev.next_trace_mask();
    F is_last_step = ev.next_trace_mask();
    F input_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    ev.add_constraint(h.is_last_idx * (h.is_last_idx - one));
    ev.add_constraint(is_last_step * (is_last_step - one));
    ev.add_constraint(next_acc_val - (acc_val + input_val));
    ev.add_constraint((out_val - next_acc_val) * is_last_step);
    reduce_transitions(ev, h);
    ev.add_to_relation(node, input_mult, input_val, h.input_id);
    ev.add_to_relation(node, out_mult, out_val, h.node_id);
    ev.finalize_logup();
}

#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_max_reduce(E& ev, const Relation2& node) {
    typedef typename E::F F;
    ReduceHead<E> h;
    read_reduce_head(ev, h);
    F input_val = ev.next_trace_mask();
    F out_val = ev.next_trace_mask();
    F max_val = ev.next_trace_mask();
    F next_max_val = ev.next_trace_mask();
    F is_last_step = ev.next_trace_mask();
    F is_max = ev.next_trace_mask();
    F input_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    ev.add_constraint(h.is_last_idx * (h.is_last_idx - one));
    ev.add_constraint(is_last_step * (is_last_step - one));
    ev.add_constraint(is_max * (is_max - one));
    ev.add_constraint(is_max * (next_max_val - input_val));
    ev.add_constraint((one - is_max) * (next_max_val - max_val));
    ev.add_constraint((out_val - next_max_val) * is_last_step);
    reduce_transitions(ev, h);
    ev.add_to_relation(node, input_mult, input_val, h.input_id);
    ev.add_to_relation(node, out_mult, out_val, h.node_id);
    ev.finalize_logup();
}

#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_contiguous(E& ev, const Relation2& node) {
    typedef typename E::F F;
    ReduceHead<E> h;
    read_reduce_head(ev, h);
    F input = ev.next_trace_mask();
    F out = ev.next_trace_mask();
    F input_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    ev.add_constraint(h.is_last_idx * (h.is_last_idx - one));
    reduce_transitions(ev, h);
    ev.add_to_relation(node, input_mult, input, h.input_id);
    ev.add_to_relation(node, out_mult, out, h.node_id);
    ev.finalize_logup();
}

// compile-time dispatch (device kernels) and run-time dispatch (host evaluators)
#pragma nv_exec_check_disable
template <int KIND, class E>
__host__ __device__ __forceinline__ void eval_kind(E& ev, const Relation2& node) {
    if (KIND == COMP_ADD) eval_add(ev, node);
    else if (KIND == COMP_MUL) eval_mul<E, false>(ev, node);
    else if (KIND == COMP_MUL_ARTIFACT) eval_mul<E, true>(ev, node);
    else if (KIND == COMP_SUM_REDUCE) eval_sum_reduce(ev, node);
    else if (KIND == COMP_MAX_REDUCE) eval_max_reduce(ev, node);
    else if (KIND == COMP_CONTIGUOUS) eval_contiguous(ev, node);
    else eval_inputs(ev, node);
}

template <class E>
__host__ __device__ __forceinline__ void eval_component(int kind, E& ev, const Relation2& node) {
    if (kind == COMP_ADD)
        eval_add(ev, node);
    else if (kind == COMP_MUL)
        eval_mul<E, false>(ev, node);
    else if (kind == COMP_MUL_ARTIFACT)
        eval_mul<E, true>(ev, node);
    else if (kind == COMP_SUM_REDUCE)
        eval_sum_reduce(ev, node);
    else if (kind == COMP_MAX_REDUCE)
        eval_max_reduce(ev, node);
    else if (kind == COMP_CONTIGUOUS)
        eval_contiguous(ev, node);
    else
        eval_inputs(ev, node);
}

// LogUp terms of the interaction-trace writers (add/witness.rs:98-104 and siblings):
// per fraction (multiplicity column, value column, id column) as main-trace column indices.
struct LookupTerm {
    int mult, val, id;
};
__host__ __device__ constexpr LookupTerm lookup_term(int kind, int k) {
    return kind == COMP_ADD   ? (k == 0 ? LookupTerm{12, 9, 1} : k == 1 ? LookupTerm{13, 10, 2} : LookupTerm{14, 11, 0})
           : (kind == COMP_MUL || kind == COMP_MUL_ARTIFACT)
               ? (k == 0 ? LookupTerm{13, 9, 1} : k == 1 ? LookupTerm{14, 10, 2} : LookupTerm{15, 11, 0})
           : kind == COMP_SUM_REDUCE ? (k == 0 ? LookupTerm{12, 7, 1} : LookupTerm{13, 8, 0})
           : kind == COMP_MAX_REDUCE ? (k == 0 ? LookupTerm{13, 7, 1} : LookupTerm{14, 8, 0})
           : kind == COMP_CONTIGUOUS ? (k == 0 ? LookupTerm{9, 7, 1} : LookupTerm{10, 8, 0})
                                     : LookupTerm{6, 5, 0};
}

}  // namespace lb
