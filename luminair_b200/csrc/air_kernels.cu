// AIR-side kernels for sm_100a: LogUp interaction-trace generation and batched constraint-quotient
// evaluation over the evaluation domain.
//
//   logup_interaction_trace  replaces `InteractionClaimGenerator::write_interaction_trace`
//       (/root/reference/crates/air/src/components/add/witness.rs:126-167 and siblings), i.e. stwo's
//       LogupTraceGenerator::{new_col, write_frac, finalize_col, finalize_last}   
//   constraint_quotients     replaces ComponentProver::evaluate_constraint_quotients_on_domain for
//       FrameworkComponent<XEval> (crates/air/src/components/mod.rs:530-601)      
//
// Both are one thread per row over column-major data (coalesced 128-byte warp reads, each input
// byte read once); the LogUp prefix sum runs in canonic-coset order over the bit-reversed
// storage as a three-phase scan.
#include "launch.cuh"
#include "kernels.cuh"

namespace lb {

// ------------------------------------------------------------------------------------
// LogUp fractions.  Column k (QM31, 4 coordinate columns) = sum_{k' <= k} mult_k' / denom_k',
// denom = alpha * id + val - z.  The n_fracs inversions of a row share one QM31 inversion.
// ------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) logup_fracs_kernel(const uint32_t* __restrict__ main, size_t main_stride, const PreCols pre_cols,
                                                          uint32_t* __restrict__ inter, size_t inter_stride, uint32_t n,
                                                          const __grid_constant__ Relations rels) {
    pdl_wait();
    constexpr int NF = component_shape(KIND).n_fracs;
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    QM31 den[NF], pre[NF];
    uint32_t mult[NF];
    QM31 run = q_from_m(1);
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        const LookupTerm t = lookup_term(KIND, k);
        const Relation2& rel = rels.r[t.rel];
        uint32_t v0 = t.pre ? pre_cols.p[t.v0][j] : main[(size_t)t.v0 * main_stride + j];
        uint32_t m = main[(size_t)t.mult * main_stride + j];
        mult[k] = t.neg ? m_neg(m) : m;
        QM31 d;
        if (t.v1 >= 0) {
            uint32_t v1 = t.pre ? pre_cols.p[t.v1][j] : main[(size_t)t.v1 * main_stride + j];
            d = q_mul_m(rel.alpha, v1);
            d.a.a = m_add(d.a.a, v0);
        } else {
            d = q_from_m(v0);
        }
        d = q_sub(d, rel.z);
        den[k] = d;
        pre[k] = run;
        run = q_mul(run, d);
    }
    QM31 inv = q_inv(run);
    QM31 dinv[NF];
#pragma unroll
    for (int k = NF - 1; k >= 0; --k) {
        dinv[k] = q_mul(inv, pre[k]);
        inv = q_mul(inv, den[k]);
    }
    QM31 acc = q_zero();
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        acc = q_add(acc, q_mul_m(dinv[k], mult[k]));
        uint32_t* o = inter + (size_t)(4 * k) * inter_stride + j;
        o[0] = acc.a.a;
        o[inter_stride] = acc.a.b;
        o[2 * inter_stride] = acc.b.a;
        o[3 * inter_stride] = acc.b.b;
    }
}

// storage index of canonic-coset point k (CanonicCoset(log).coset order):
//   circle-domain index = k/2 (k even) | n - (k+1)/2 (k odd);  storage = bitrev(domain index)
__device__ __forceinline__ uint32_t coset_to_storage(uint32_t k, int log) {
    uint32_t n = 1u << log;
    uint32_t d = (k & 1) ? (n - ((k + 1) >> 1)) : (k >> 1);
    return log ? (__brev(d) >> (32 - log)) : 0;
}

constexpr int SCAN_BLOCK = 1024;  // elements per CTA (256 threads x 4)

// phase 1: per-CTA inclusive scan of the last column in coset order (gathered), mod P
__global__ void __launch_bounds__(256) logup_scan_local_kernel(const uint32_t* __restrict__ col, size_t coord_stride,
                                                               uint32_t* __restrict__ tmp, uint32_t* __restrict__ block_sums,
                                                               int log) {
    pdl_wait();
    __shared__ uint32_t s_warp[8];
    const uint32_t n = 1u << log;
    const int coord = blockIdx.y;
    const uint32_t* src = col + (size_t)coord * coord_stride;
    uint32_t* dst = tmp + (size_t)coord * n;
    const uint32_t k0 = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (k0 + i < n) ? src[coset_to_storage(k0 + i, log)] : 0u;
    v[1] = m_add(v[1], v[0]);
    v[2] = m_add(v[2], v[1]);
    v[3] = m_add(v[3], v[2]);
    uint32_t tot = v[3];
    // warp inclusive scan of the per-thread totals
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t s = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if (lane >= o) s = m_add(s, t);
    }
    if (lane == 31) s_warp[w] = s;
    __syncthreads();
    uint32_t woff = 0;
    for (int i = 0; i < w; ++i) woff = m_add(woff, s_warp[i]);
    uint32_t excl = m_add(woff, m_sub(s, tot));
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (k0 + i < n) dst[k0 + i] = m_add(v[i], excl);
    if (threadIdx.x == 255) block_sums[(size_t)coord * gridDim.x + blockIdx.x] = m_add(s, woff);
}

// phase 2: exclusive scan of the CTA totals (<= 2^24 / 1024 entries per coordinate); total -> claimed.
// One CTA per coordinate: every thread scans a contiguous chunk, the 256 chunk totals are scanned in shared memory.
__global__ void __launch_bounds__(256) logup_scan_sums_kernel(uint32_t* block_sums, uint32_t n_blocks, uint32_t* claimed) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ uint32_t s_tot[256];
    const int coord = blockIdx.x;
    uint32_t* s = block_sums + (size_t)coord * n_blocks;
    const uint32_t per = (n_blocks + 255) / 256;
    const uint32_t b0 = min(threadIdx.x * per, n_blocks), b1 = min(b0 + per, n_blocks);
    uint32_t run = 0;
    for (uint32_t b = b0; b < b1; ++b) run = m_add(run, s[b]);
    s_tot[threadIdx.x] = run;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int t = 0; t < 256; ++t) {
            uint32_t v = s_tot[t];
            s_tot[t] = acc;
            acc = m_add(acc, v);
        }
        claimed[coord] = acc;
    }
    __syncthreads();
    run = s_tot[threadIdx.x];
    for (uint32_t b = b0; b < b1; ++b) {
        uint32_t t = s[b];
        s[b] = run;
        run = m_add(run, t);
    }
}

// phase 3: out[storage(k)] = local[k] + offset[block] - (k + 1) * claimed / n
__global__ void __launch_bounds__(256) logup_scan_apply_kernel(uint32_t* __restrict__ col, size_t coord_stride,
                                                               const uint32_t* __restrict__ tmp,
                                                               const uint32_t* __restrict__ block_sums,
                                                               const uint32_t* __restrict__ claimed, int log, uint32_t inv_n) {
    pdl_wait();
    const uint32_t n = 1u << log;
    const int coord = blockIdx.y;
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t n_blocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    uint32_t shift = m_mul(claimed[coord], inv_n);
    uint32_t v = m_add(tmp[(size_t)coord * n + k], block_sums[(size_t)coord * n_blocks + k / SCAN_BLOCK]);
    v = m_sub(v, m_mul(shift, (k + 1) % P));
    col[(size_t)coord * coord_stride + coset_to_storage(k, log)] = v;
}

namespace {
struct LogupLaunch {
    const uint32_t* main;
    size_t main_stride;
    PreCols pre;
    uint32_t* inter;
    size_t inter_stride;
    uint32_t n;
    const Relations& rels;
    cudaStream_t stream;
    template <int KIND>
    void operator()() {
        // the artifact-era Mul shares the LogUp terms of Mul
        constexpr int K = KIND == COMP_MUL_ARTIFACT ? COMP_MUL : KIND;
        launch_k(logup_fracs_kernel<K>, (n + 255) / 256, 256, 0, stream, main, main_stride, pre, inter, inter_stride, n, rels);
    }
};
}  // namespace

cudaError_t logup_interaction_trace(int kind, const uint32_t* main, size_t main_stride, PreCols pre, uint32_t* inter,
                                    size_t inter_stride, int log, const Relations& rels, uint32_t* d_scan_tmp,
                                    uint32_t* d_block_sums, uint32_t* d_claimed, cudaStream_t stream) {
    uint32_t n = 1u << log;
    unsigned blocks = (n + 255) / 256;
    if (kind < 0 || kind >= COMP_KIND_COUNT) return cudaErrorInvalidValue;
    int nf = component_shape(kind).n_fracs;
    if (component_shape(kind).n_pre > 0 && !pre.p[0]) return cudaErrorInvalidValue;
    LogupLaunch launch{main, main_stride, pre, inter, inter_stride, n, rels, stream};
    if (!dispatch_kind(kind, launch)) return cudaErrorInvalidValue;
    uint32_t* last = inter + (size_t)(4 * (nf - 1)) * inter_stride;
    uint32_t n_blocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    launch_k(logup_scan_local_kernel, dim3(n_blocks, 4), 256, 0, stream, last, inter_stride, d_scan_tmp, d_block_sums, log);
    launch_k(logup_scan_sums_kernel, 4, 256, 0, stream, d_block_sums, n_blocks, d_claimed);
    uint32_t inv_n = m_inv(n % P);
    launch_k(logup_scan_apply_kernel, dim3(blocks, 4), 256, 0, stream, last, inter_stride, d_scan_tmp, d_block_sums, d_claimed, log,
                                                                 inv_n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// Constraint quotients: one thread per evaluation-domain row.
// ------------------------------------------------------------------------------------
struct DomainEval : LogupMixin<DomainEval, FM, FQ> {
    typedef FM F;
    typedef FQ EF;
    const ConstraintParams& p;
    uint32_t row, row_prev;
    int mc = 0, ic = 0, ci = 0;
    QM31 res;
    QM31 cumsum_shift;

    __device__ __forceinline__ DomainEval(const ConstraintParams& p_, uint32_t row_, uint32_t row_prev_)
        : p(p_), row(row_), row_prev(row_prev_), res(q_zero()), cumsum_shift(p_.cumsum_shift) {}

    __device__ __forceinline__ FM constant(uint32_t c) const { return {c}; }
    __device__ __forceinline__ FM get_preprocessed_column(int k) const { return {p.pre.p[k][row]}; }
    __device__ __forceinline__ FM next_trace_mask() {
        FM v = {p.main[(size_t)mc * p.main_stride + row]};
        ++mc;
        return v;
    }
    __device__ __forceinline__ FQ read_ext(uint32_t r) const {
        const uint32_t* b = p.inter + (size_t)ic * p.inter_stride + r;
        return {q_make(b[0], b[p.inter_stride], b[2 * p.inter_stride], b[3 * p.inter_stride])};
    }
    __device__ __forceinline__ FQ next_ext_mask_cur() {
        FQ v = read_ext(row);
        ic += 4;
        return v;
    }
    __device__ __forceinline__ void next_ext_mask_prev_cur(FQ& prev, FQ& cur) {
        if (p.inter_prev) {
            const uint32_t* b = p.inter_prev + row;
            prev = {q_make(b[0], b[p.inter_stride], b[2 * p.inter_stride], b[3 * p.inter_stride])};
        } else {
            prev = read_ext(row_prev);
        }
        cur = read_ext(row);
        ic += 4;
    }
    __device__ __forceinline__ void add_constraint(FM c) {
        res = q_add(res, q_mul_m(p.pows[ci], c.v));
        ++ci;
    }
    __device__ __forceinline__ void add_constraint_ef(FQ c) {
        res = q_add(res, q_mul(p.pows[ci], c.v));
        ++ci;
    }
};

// offset_bit_reversed_circle_domain_index(row, domain_log, eval_log, -1)
__device__ __forceinline__ uint32_t prev_row_index(uint32_t j, int domain_log, int eval_log) {
    uint32_t n = 1u << eval_log, half = n >> 1;
    uint32_t nat = __brev(j) >> (32 - eval_log);
    uint32_t step = 1u << (eval_log - domain_log - 1);  // |offset| = 1
    uint32_t r;
    if (nat < half)
        r = (nat + half - step) & (half - 1);
    else
        r = ((nat - half + step) & (half - 1)) + half;
    return __brev(r) >> (32 - eval_log);
}

template <int KIND>
__global__ void __launch_bounds__(256) constraint_quotients_kernel(const __grid_constant__ ConstraintParams p) {
    pdl_wait();
    uint32_t n = p.n_rows ? p.n_rows : (1u << p.eval_log);
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;  // local row; the global row is row0 + j
    if (j >= n) return;
    DomainEval ev(p, j, p.inter_prev ? 0u : prev_row_index(j, p.log_size, p.eval_log));
    eval_kind<KIND>(ev, p.rels);
    QM31 r = q_mul_m(ev.res, p.denom_inv[(p.row0 + j) >> p.log_size]);
    if (p.accumulate) r = q_add(r, q_make(p.acc[0][j], p.acc[1][j], p.acc[2][j], p.acc[3][j]));
    p.acc[0][j] = r.a.a;
    p.acc[1][j] = r.a.b;
    p.acc[2][j] = r.b.a;
    p.acc[3][j] = r.b.b;
}

namespace {
struct ConstraintLaunch {
    const ConstraintParams& p;
    cudaStream_t stream;
    template <int KIND>
    void operator()() {
        uint32_t n = p.n_rows ? p.n_rows : (1u << p.eval_log);
        launch_k(constraint_quotients_kernel<KIND>, (n + 255) / 256, 256, 0, stream, p);
    }
};
}  // namespace

// out[j] = col[prev(j)], j < 2^eval_log, prev = offset_bit_reversed_circle_domain_index(., domain_log, eval_log, -1); source
// and destination are each held as 2^logw equal row ranges (logw = 0: one plain column)
struct ShardPtrs {
    const uint32_t* p[8];
};
struct ShardOut {
    uint32_t* p[8];
};
__global__ void __launch_bounds__(256) shifted_prev_kernel(ShardOut out, ShardPtrs col, int src_shard_log, int dst_shard_log,
                                                           int domain_log, int eval_log) {
    pdl_wait();
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (1u << eval_log)) return;
    uint32_t q = prev_row_index(j, domain_log, eval_log);
    out.p[j >> dst_shard_log][j & ((1u << dst_shard_log) - 1)] = col.p[q >> src_shard_log][q & ((1u << src_shard_log) - 1)];
}
cudaError_t shifted_prev_column(uint32_t* const out_shards[8], int n_out, const uint32_t* const src_shards[8], int n_src,
                                int domain_log, int eval_log, cudaStream_t stream) {
    auto lg = [](int n) {
        int l = 0;
        while ((1 << l) < n) ++l;
        return l;
    };
    if (eval_log - domain_log < 1 || eval_log > 30 || n_out < 1 || n_out > 8 || n_src < 1 || n_src > 8 || (n_out & (n_out - 1)) ||
        (n_src & (n_src - 1)))
        return cudaErrorInvalidValue;
    ShardPtrs sp{};
    ShardOut so{};
    for (int k = 0; k < n_src; ++k) sp.p[k] = src_shards[k];
    for (int k = 0; k < n_out; ++k) so.p[k] = out_shards[k];
    uint32_t n = 1u << eval_log;
    launch_k(shifted_prev_kernel, (n + 255) / 256, 256, 0, stream, so, sp, eval_log - lg(n_src), eval_log - lg(n_out), domain_log, eval_log);
    return cudaGetLastError();
}

cudaError_t constraint_quotients(int kind, const ConstraintParams& p, cudaStream_t stream) {
    if (p.eval_log - p.log_size < 1 || p.eval_log > 30 || !p.denom_inv) return cudaErrorInvalidValue;
    if (p.n_rows && ((uint64_t)p.row0 + p.n_rows > ((uint64_t)1 << p.eval_log) || !p.inter_prev)) return cudaErrorInvalidValue;
    if (kind < 0 || kind >= COMP_KIND_COUNT) return cudaErrorInvalidValue;
    if (component_shape(kind).n_pre > 0 && !p.pre.p[0]) return cudaErrorInvalidValue;
    ConstraintLaunch launch{p, stream};
    if (!dispatch_kind(kind, launch)) return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace lb
