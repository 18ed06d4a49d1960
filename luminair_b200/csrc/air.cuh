// LuminAIR component AIRs written once against an abstract row evaluator, the way the reference
// writes `FrameworkEval::evaluate<E: EvalAtRow>`:
//   Add     /root/reference/crates/air/src/components/add/component.rs:38-116
//   Mul     /root/reference/crates/air/src/components/mul/component.rs:40-128
//   Inputs  /root/reference/crates/air/src/components/inputs/components.rs:37-85
//   SumReduce / MaxReduce / Contiguous, Recip / Sqrt / Rem, Sin / Exp2 / Log2 (+ their lookup-table components),
//   LessThan (+ RangeCheckLookup): crates/air/src/components/<name>/component.rs, cited at each template
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

// LuminairInteractionElements (components/mod.rs:220-236, lookups/mod.rs:31-51): NodeElements, then the LUT
// relations sin, exp2, log2 (2 values each) and range_check (1 value).
enum RelationId { REL_NODE = 0, REL_SIN = 1, REL_EXP2 = 2, REL_LOG2 = 3, REL_RANGE_CHECK = 4, REL_COUNT = 5 };
struct Relations {
    Relation2 r[REL_COUNT];
};

constexpr uint32_t FP_SCALE = 1u << 12;  // DEFAULT_FP_SCALE, crates/air/src/lib.rs:23
constexpr int MAX_FRACS = 7;             // less_than: 3 node + 4 range-check relation uses
constexpr int MAX_CONSTRAINTS = 20;
constexpr int MAX_MAIN_COLS = 24;

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
    // relation!(X, 1): combine(v) = v - z
#pragma nv_exec_check_disable
    __host__ __device__ __forceinline__ void add_to_relation1(const Relation2& rel, F multiplicity, F v0) {
        EF d = (EF{q_zero()} + v0) - EF{rel.z};
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
    COMP_RECIP = 7,       // components/recip/component.rs:39-107
    COMP_SQRT = 8,        // components/sqrt/component.rs:38-107
    COMP_REM = 9,         // components/rem/component.rs:38-124
    COMP_SIN = 10,        // components/sin/component.rs:51-123
    COMP_EXP2 = 11,       // components/exp2/component.rs:46-118
    COMP_LOG2 = 12,       // components/log2/component.rs:46-117
    COMP_SIN_LOOKUP = 13,   // components/lookups/sin/component.rs:41-59
    COMP_EXP2_LOOKUP = 14,  // components/lookups/exp2/component.rs:41-59
    COMP_LOG2_LOOKUP = 15,  // components/lookups/log2/component.rs:40-58
    COMP_LESS_THAN = 16,    // components/less_than/component.rs:49-184
    COMP_RANGE_CHECK_LOOKUP = 17,  // components/lookups/range_check/component.rs:44-60
    COMP_KIND_COUNT = 18
};

struct ComponentShape {
    int n_main;         // main-trace columns (add/witness.rs:24 etc.)
    int n_fracs;        // LogUp relation uses -> n_fracs QM31 interaction columns
    int n_constraints;  // counted by InfoEval at start-up, checked against this table
    int padding_one_col;  // index of `is_last_idx` (1 in the padding row); -1: none
    int n_pre;          // preprocessed (LUT) columns the component reads
    int lut;            // RelationId of the LUT this component consumes or tabulates (0: none)
};
__host__ __device__ constexpr ComponentShape component_shape(int kind) {
    return kind == COMP_ADD            ? ComponentShape{15, 3, 9, 4, 0, 0}
           : kind == COMP_MUL          ? ComponentShape{16, 3, 9, 4, 0, 0}
           : kind == COMP_MUL_ARTIFACT ? ComponentShape{16, 3, 10, 4, 0, 0}
           : kind == COMP_SUM_REDUCE   ? ComponentShape{14, 2, 9, 3, 0, 0}
           : kind == COMP_MAX_REDUCE   ? ComponentShape{15, 2, 11, 3, 0, 0}
           : kind == COMP_CONTIGUOUS   ? ComponentShape{11, 2, 6, 3, 0, 0}
           : kind == COMP_RECIP        ? ComponentShape{13, 2, 7, 3, 0, 0}
           : kind == COMP_SQRT         ? ComponentShape{13, 2, 7, 3, 0, 0}
           : kind == COMP_REM          ? ComponentShape{16, 3, 9, 4, 0, 0}
           : kind == COMP_SIN          ? ComponentShape{12, 3, 7, 3, 0, REL_SIN}
           : kind == COMP_EXP2         ? ComponentShape{12, 3, 7, 3, 0, REL_EXP2}
           : kind == COMP_LOG2         ? ComponentShape{12, 3, 7, 3, 0, REL_LOG2}
           : kind == COMP_SIN_LOOKUP   ? ComponentShape{1, 1, 1, -1, 2, REL_SIN}
           : kind == COMP_EXP2_LOOKUP  ? ComponentShape{1, 1, 1, -1, 2, REL_EXP2}
           : kind == COMP_LOG2_LOOKUP  ? ComponentShape{1, 1, 1, -1, 2, REL_LOG2}
           : kind == COMP_LESS_THAN    ? ComponentShape{22, 7, 16, 4, 0, REL_RANGE_CHECK}
           : kind == COMP_RANGE_CHECK_LOOKUP ? ComponentShape{1, 1, 1, -1, 1, REL_RANGE_CHECK}
                                       : ComponentShape{7, 1, 4, 2, 0, 0};
}
// `padding()` of each *TraceTableRow (e.g. add/table.rs): is_last_idx = 1, everything else 0, except
// less_than/table.rs padding(): the row "0 < 1" (rhs = 1, out = 1.0, diff = 1, limb0 = 1).
__host__ __device__ constexpr uint32_t padding_value(int kind, int col) {
    return col == component_shape(kind).padding_one_col ? 1u
           : kind != COMP_LESS_THAN                     ? 0u
           : (col == 10 || col == 12 || col == 14)      ? 1u
           : col == 11                                  ? FP_SCALE
                                                        : 0u;
}
// a LUT consumer is evaluated on CanonicCoset(max(log_size, lut_log_size) + 1) (exp2/component.rs:41-43,
// less_than/component.rs:44-46); everything else on CanonicCoset(log_size + 1)
__host__ __device__ constexpr bool consumes_lut(int kind) {
    return kind == COMP_SIN || kind == COMP_EXP2 || kind == COMP_LOG2 || kind == COMP_LESS_THAN;
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


// numerair EvalFixedPoint::{eval_fixed_recip, eval_fixed_sqrt, eval_fixed_rem} (un-vendored; not covered by the
// reference's committed proof -> "parity unpinned"): the identities the operators emit rows for
// (crates/graph/src/op/prim.rs:375, :604, :1359) in the same form as the pinned eval_fixed_mul.
#pragma nv_exec_check_disable
template <class E, bool SQRT>
__host__ __device__ __forceinline__ void eval_recip_sqrt(E& ev, const Relation2& node) {
    typedef typename E::F F;
    ReduceHead<E> h;
    read_reduce_head(ev, h);
    F input_val = ev.next_trace_mask();
    F out_val = ev.next_trace_mask();
    F rem_val = ev.next_trace_mask();
    F scale = ev.next_trace_mask();
    F input_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    ev.add_constraint(h.is_last_idx * (h.is_last_idx - one));
    if (SQRT)
        ev.add_constraint(input_val * scale - (out_val * out_val + rem_val));  // eval_fixed_sqrt
    else
        ev.add_constraint(scale * scale - (input_val * out_val + rem_val));  // eval_fixed_recip
    reduce_transitions(ev, h);
    ev.add_to_relation(node, input_mult, input_val, h.input_id);
    ev.add_to_relation(node, out_mult, out_val, h.node_id);
    ev.finalize_logup();
}

#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_rem(E& ev, const Relation2& node) {
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
    F rem_val = ev.next_trace_mask();
    F quotient = ev.next_trace_mask();
    F lhs_mult = ev.next_trace_mask();
    F rhs_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    ev.add_constraint(is_last_idx * (is_last_idx - one));
    ev.add_constraint(lhs_val - (quotient * rhs_val + rem_val));  // eval_fixed_rem
    F not_last = one - is_last_idx;
    ev.add_constraint(not_last * (next_node_id - node_id));
    ev.add_constraint(not_last * (next_lhs_id - lhs_id));
    ev.add_constraint(not_last * (next_rhs_id - rhs_id));
    ev.add_constraint(not_last * (next_idx - idx - one));
    ev.add_to_relation(node, lhs_mult, lhs_val, lhs_id);
    ev.add_to_relation(node, rhs_mult, rhs_val, rhs_id);
    ev.add_to_relation(node, out_mult, rem_val, node_id);
    ev.finalize_logup();
}

// Sin / Exp2 / Log2: the (input, output) pair is looked up in the function's table
#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_lut_consumer(E& ev, const Relation2& node, const Relation2& lut) {
    typedef typename E::F F;
    ReduceHead<E> h;
    read_reduce_head(ev, h);
    F input_val = ev.next_trace_mask();
    F out_val = ev.next_trace_mask();
    F input_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F lookup_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    ev.add_constraint(h.is_last_idx * (h.is_last_idx - one));
    reduce_transitions(ev, h);
    ev.add_to_relation(node, input_mult, input_val, h.input_id);
    ev.add_to_relation(node, out_mult, out_val, h.node_id);
    ev.add_to_relation(lut, lookup_mult, input_val, out_val);
    ev.finalize_logup();
}

// SinLookup / Exp2Lookup / Log2Lookup (N_PRE = 2) and RangeCheckLookup (N_PRE = 1): every table row yields
// -multiplicity
#pragma nv_exec_check_disable
template <class E, int N_PRE>
__host__ __device__ __forceinline__ void eval_lut_table(E& ev, const Relation2& lut) {
    typedef typename E::F F;
    F lut0 = ev.get_preprocessed_column(0);
    F lut1 = N_PRE == 2 ? ev.get_preprocessed_column(1) : lut0;
    F multiplicity = ev.next_trace_mask();
    F neg = ev.constant(0) - multiplicity;
    if (N_PRE == 2)
        ev.add_to_relation(lut, neg, lut0, lut1);
    else
        ev.add_to_relation1(lut, neg, lut0);
    ev.finalize_logup();
}

#pragma nv_exec_check_disable
template <class E>
__host__ __device__ __forceinline__ void eval_less_than(E& ev, const Relation2& node, const Relation2& range_check) {
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
    F diff_val = ev.next_trace_mask();
    F borrow = ev.next_trace_mask();
    F limb0 = ev.next_trace_mask();
    F limb1 = ev.next_trace_mask();
    F limb2 = ev.next_trace_mask();
    F limb3 = ev.next_trace_mask();
    F lhs_mult = ev.next_trace_mask();
    F rhs_mult = ev.next_trace_mask();
    F out_mult = ev.next_trace_mask();
    F diff_mult = ev.next_trace_mask();
    F one = ev.constant(1);
    // the reference passes TWO_POW_31_MINUS_1 = 2^31 - 1 = p, i.e. the zero element, as "2^k"
    // (less_than/component.rs:51, crates/air/src/lib.rs:26)
    F two_pow_k = ev.constant(0);
    F scale_factor = ev.constant(FP_SCALE);
    ev.add_constraint(is_last_idx * (is_last_idx - one));
    ev.add_constraint(borrow * (borrow - one));
    ev.add_constraint(out_val - ((one - borrow) * scale_factor));
    ev.add_constraint(lhs_val + diff_val - rhs_val - (borrow * two_pow_k));
    F recomposed = limb3 * ev.constant(1u << 24) + limb2 * ev.constant(1u << 16) + limb1 * ev.constant(1u << 8) + limb0;
    ev.add_constraint(diff_val - recomposed);
    F not_last = one - is_last_idx;
    ev.add_constraint(not_last * (next_node_id - node_id));
    ev.add_constraint(not_last * (next_lhs_id - lhs_id));
    ev.add_constraint(not_last * (next_rhs_id - rhs_id));
    ev.add_constraint(not_last * (next_idx - idx - one));
    ev.add_to_relation(node, lhs_mult, lhs_val, lhs_id);
    ev.add_to_relation(node, rhs_mult, rhs_val, rhs_id);
    ev.add_to_relation(node, out_mult, out_val, node_id);
    ev.add_to_relation1(range_check, diff_mult, limb0);
    ev.add_to_relation1(range_check, diff_mult, limb1);
    ev.add_to_relation1(range_check, diff_mult, limb2);
    ev.add_to_relation1(range_check, diff_mult, limb3);
    ev.finalize_logup();
}

// compile-time dispatch (device kernels) and run-time dispatch (host evaluators)
#pragma nv_exec_check_disable
template <int KIND, class E>
__host__ __device__ __forceinline__ void eval_kind(E& ev, const Relations& rels) {
    const Relation2& node = rels.r[REL_NODE];
    if (KIND == COMP_ADD) eval_add(ev, node);
    else if (KIND == COMP_MUL) eval_mul<E, false>(ev, node);
    else if (KIND == COMP_MUL_ARTIFACT) eval_mul<E, true>(ev, node);
    else if (KIND == COMP_SUM_REDUCE) eval_sum_reduce(ev, node);
    else if (KIND == COMP_MAX_REDUCE) eval_max_reduce(ev, node);
    else if (KIND == COMP_CONTIGUOUS) eval_contiguous(ev, node);
    else if (KIND == COMP_RECIP) eval_recip_sqrt<E, false>(ev, node);
    else if (KIND == COMP_SQRT) eval_recip_sqrt<E, true>(ev, node);
    else if (KIND == COMP_REM) eval_rem(ev, node);
    else if (KIND == COMP_SIN) eval_lut_consumer(ev, node, rels.r[REL_SIN]);
    else if (KIND == COMP_EXP2) eval_lut_consumer(ev, node, rels.r[REL_EXP2]);
    else if (KIND == COMP_LOG2) eval_lut_consumer(ev, node, rels.r[REL_LOG2]);
    else if (KIND == COMP_SIN_LOOKUP) eval_lut_table<E, 2>(ev, rels.r[REL_SIN]);
    else if (KIND == COMP_EXP2_LOOKUP) eval_lut_table<E, 2>(ev, rels.r[REL_EXP2]);
    else if (KIND == COMP_LOG2_LOOKUP) eval_lut_table<E, 2>(ev, rels.r[REL_LOG2]);
    else if (KIND == COMP_LESS_THAN) eval_less_than(ev, node, rels.r[REL_RANGE_CHECK]);
    else if (KIND == COMP_RANGE_CHECK_LOOKUP) eval_lut_table<E, 1>(ev, rels.r[REL_RANGE_CHECK]);
    else eval_inputs(ev, node);
}

// calls f.template operator()<KIND>() for the run-time `kind`; false if the kind is unknown
template <class Fn>
__host__ inline bool dispatch_kind(int kind, Fn&& f) {
    switch (kind) {
#define LB_KIND_CASE(K) case K: f.template operator()<K>(); return true;
        LB_KIND_CASE(COMP_ADD) LB_KIND_CASE(COMP_MUL) LB_KIND_CASE(COMP_INPUTS) LB_KIND_CASE(COMP_MUL_ARTIFACT)
        LB_KIND_CASE(COMP_SUM_REDUCE) LB_KIND_CASE(COMP_MAX_REDUCE) LB_KIND_CASE(COMP_CONTIGUOUS) LB_KIND_CASE(COMP_RECIP)
        LB_KIND_CASE(COMP_SQRT) LB_KIND_CASE(COMP_REM) LB_KIND_CASE(COMP_SIN) LB_KIND_CASE(COMP_EXP2) LB_KIND_CASE(COMP_LOG2)
        LB_KIND_CASE(COMP_SIN_LOOKUP) LB_KIND_CASE(COMP_EXP2_LOOKUP) LB_KIND_CASE(COMP_LOG2_LOOKUP) LB_KIND_CASE(COMP_LESS_THAN)
        LB_KIND_CASE(COMP_RANGE_CHECK_LOOKUP)
#undef LB_KIND_CASE
        default: return false;
    }
}

template <class E>
struct EvalComponentFn {
    E& ev;
    const Relations& rels;
    template <int KIND>
    void operator()() { eval_kind<KIND>(ev, rels); }
};
template <class E>
inline void eval_component(int kind, E& ev, const Relations& rels) {
    EvalComponentFn<E> f{ev, rels};
    dispatch_kind(kind, f);
}

// LogUp terms of the interaction-trace writers (add/witness.rs:98-104 and siblings, exp2/witness.rs:119-160,
// less_than/witness.rs:144-226, lookups/exp2/witness.rs:117-144): per fraction the multiplicity column and the
// combined values.  v0, v1 index main-trace columns, or preprocessed columns when `pre`; v1 < 0: one-value
// relation; `neg`: the numerator is -multiplicity (table side of a LUT).
struct LookupTerm {
    int mult, v0, v1;
    int rel = REL_NODE;
    bool neg = false, pre = false;
};
__host__ __device__ constexpr LookupTerm lookup_term(int kind, int k) {
    const int lut = component_shape(kind).lut;
    return kind == COMP_ADD   ? (k == 0 ? LookupTerm{12, 9, 1} : k == 1 ? LookupTerm{13, 10, 2} : LookupTerm{14, 11, 0})
           : (kind == COMP_MUL || kind == COMP_MUL_ARTIFACT)
               ? (k == 0 ? LookupTerm{13, 9, 1} : k == 1 ? LookupTerm{14, 10, 2} : LookupTerm{15, 11, 0})
           : kind == COMP_SUM_REDUCE ? (k == 0 ? LookupTerm{12, 7, 1} : LookupTerm{13, 8, 0})
           : kind == COMP_MAX_REDUCE ? (k == 0 ? LookupTerm{13, 7, 1} : LookupTerm{14, 8, 0})
           : kind == COMP_CONTIGUOUS ? (k == 0 ? LookupTerm{9, 7, 1} : LookupTerm{10, 8, 0})
           : (kind == COMP_RECIP || kind == COMP_SQRT) ? (k == 0 ? LookupTerm{11, 7, 1} : LookupTerm{12, 8, 0})
           : kind == COMP_REM ? (k == 0 ? LookupTerm{13, 9, 1} : k == 1 ? LookupTerm{14, 10, 2} : LookupTerm{15, 11, 0})
           : (kind == COMP_SIN || kind == COMP_EXP2 || kind == COMP_LOG2)
               ? (k == 0 ? LookupTerm{9, 7, 1} : k == 1 ? LookupTerm{10, 8, 0} : LookupTerm{11, 7, 8, lut})
           : (kind == COMP_SIN_LOOKUP || kind == COMP_EXP2_LOOKUP || kind == COMP_LOG2_LOOKUP) ? LookupTerm{0, 0, 1, lut, true, true}
           : kind == COMP_RANGE_CHECK_LOOKUP ? LookupTerm{0, 0, -1, lut, true, true}
           : kind == COMP_LESS_THAN
               ? (k == 0 ? LookupTerm{18, 9, 1} : k == 1 ? LookupTerm{19, 10, 2} : k == 2 ? LookupTerm{20, 11, 0}
                                                                                          : LookupTerm{21, 11 + k, -1, lut})
               : LookupTerm{6, 5, 0};
}

}  // namespace lb
