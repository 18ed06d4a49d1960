// gen_trace on the device: the `process_trace` loops of every LuminAIR operator
// (crates/graph/src/op/prim.rs) as CUDA kernels.  A tensor is a device array of raw Fixed<12> values (int32);
// an operand is a tensor plus an optional gather index (the operator's index expression for broadcasts / expands,
// already flattened by the caller).  Each kernel computes the node's output tensor and appends the node's rows, in the
// column order of the component's *TraceTableRow (crates/air/src/components/*/table.rs), to a row-major device table -
// the format lb_prove takes with rows_on_device = 1.  One thread per row (per output element for the reductions, whose
// rows are a sequential recurrence); a CTA stages its run of rows in shared memory and stores it coalesced.
//
//   CopyToStwo / LuminairConstant  prim.rs:52-84, 151-188     LuminairContiguous  prim.rs:229-298
//   LuminairRecip  :388-428   LuminairSin  :496-543   LuminairSqrt  :617-657   LuminairExp2  :725-772
//   LuminairLog2   :840-887   LuminairAdd  :967-1013  LuminairMul   :1090-1136 LuminairLessThan :1225-1292
//   LuminairRem    :1372-1418 LuminairSumReduce :1517-1562   LuminairMaxReduce :1685-1731
//
// Lookup operators do not call libm on the device: the caller generated the LUT columns on the host exactly as the
// reference does (preprocessed.rs:351-383), so `f(x)` is read from LUT column 1 at `find_index(x)` (preprocessed.rs:96-116)
// - bit-identical to the reference's f64 path by construction - and the entry's multiplicity counter is bumped with one
// atomic (the *_lookup table's only column).  LessThan bumps the 8-bit range-check counters of its four limbs through a
// per-CTA shared-memory histogram.
#include "launch.cuh"
#include "kernels.cuh"

namespace lb {
namespace {

__device__ __forceinline__ uint32_t to_m31(int64_t v) {  // Fixed::to_m31: negative raw values map to p - |v|
    return v < 0 ? (uint32_t)((int64_t)P + v) : (uint32_t)v;
}
__device__ __forceinline__ int64_t rd(const int32_t* __restrict__ v, const uint32_t* __restrict__ idx, uint64_t i) {
    return v[idx ? (uint64_t)idx[i] : i];
}
__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) {
    int64_t q = a / b;
    return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}
__device__ __forceinline__ int64_t isqrt_floor(int64_t t) {
    if (t <= 0) return 0;
    int64_t s = (int64_t)sqrt((double)t);
    while (s * s > t) --s;
    while ((s + 1) * (s + 1) <= t) ++s;
    return s;
}

// binary head: node, lhs, rhs, idx, is_last_idx, next_node, next_lhs, next_rhs, next_idx
__device__ __forceinline__ uint32_t* head2(uint32_t* r, const TraceOp& p, uint64_t i) {
    r[0] = p.node_id; r[1] = p.lhs_id; r[2] = p.rhs_id; r[3] = (uint32_t)i; r[4] = (i == p.n - 1);
    r[5] = p.node_id; r[6] = p.lhs_id; r[7] = p.rhs_id; r[8] = (uint32_t)(i + 1);
    return r + 9;
}
// unary head: node, input, idx, is_last_idx, next_node, next_input, next_idx
__device__ __forceinline__ uint32_t* head1(uint32_t* r, const TraceOp& p, uint64_t i) {
    r[0] = p.node_id; r[1] = p.lhs_id; r[2] = (uint32_t)i; r[3] = (i == p.n - 1);
    r[4] = p.node_id; r[5] = p.lhs_id; r[6] = (uint32_t)(i + 1);
    return r + 7;
}

template <int OP>
__host__ __device__ constexpr int op_cols() {  // columns of the component's *TraceTableRow (crates/air/src/components/*/table.rs)
    return OP == LB_OP_INPUTS ? 7 : OP == LB_OP_ADD ? 15 : (OP == LB_OP_MUL || OP == LB_OP_REM) ? 16 : OP == LB_OP_LESS_THAN ? 22
           : (OP == LB_OP_RECIP || OP == LB_OP_SQRT) ? 13 : OP == LB_OP_CONTIGUOUS ? 11 : 12;
}

// A thread's row is NC words at a stride of NC words in the table: stored straight from registers, a warp instruction would
// touch 32 different sectors.  The CTA's 256 rows are one contiguous run of 256 * NC words, so each thread parks its row in
// shared memory (odd pitch: conflict-free) and the CTA copies the run out with 128-bit, fully coalesced stores.
template <int OP>
__global__ void __launch_bounds__(256) trace_rows_kernel(const TraceOp p) {
    pdl_wait();
    constexpr int NC = op_cols<OP>(), NCP = NC | 1;
    __shared__ uint32_t stage[256 * NCP];
    __shared__ uint32_t hist[OP == LB_OP_LESS_THAN ? 256 : 1];
    if (OP == LB_OP_LESS_THAN) {
        hist[threadIdx.x] = 0;
        __syncthreads();
    }
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t* const row = stage + threadIdx.x * NCP;
    if (i < p.n) {
        const uint32_t mult = p.out_mult ? p.out_mult[i] : p.out_mult_all;
        const uint32_t M1 = P - 1;  // multiplicity -1 of a consumed operand
        if (OP == LB_OP_INPUTS) {
            const int64_t v = p.lhs[i];
            uint32_t* r = row;  // inputs/table.rs
            r[0] = p.node_id; r[1] = (uint32_t)i; r[2] = (i == p.n - 1); r[3] = p.node_id; r[4] = (uint32_t)(i + 1);
            r[5] = to_m31(v); r[6] = mult;
        } else if (OP == LB_OP_ADD || OP == LB_OP_MUL || OP == LB_OP_REM) {
            const int64_t x = rd(p.lhs, p.lhs_idx, i), y = rd(p.rhs, p.rhs_idx, i);
            int64_t o, aux = 0;
            if (OP == LB_OP_ADD) {
                o = x + y;
            } else if (OP == LB_OP_MUL) {  // floor division by the scale, remainder in [0, 4096)
                const int64_t prod = x * y;
                o = prod >> 12;
                aux = prod - (o << 12);
            } else {  // rem: floor quotient, remainder takes the divisor's sign
                aux = y ? floor_div(x, y) : 0;
                o = x - aux * y;
            }
            p.out[i] = (int32_t)o;
            uint32_t* r = head2(row, p, i);  // add/table.rs, mul/table.rs, rem/table.rs
            *r++ = to_m31(x); *r++ = to_m31(y); *r++ = to_m31(o);
            if (OP != LB_OP_ADD) *r++ = to_m31(aux);
            *r++ = M1; *r++ = M1; *r = mult;
        } else if (OP == LB_OP_LESS_THAN) {  // less_than/table.rs: 22 columns
            const int64_t x = rd(p.lhs, p.lhs_idx, i), y = rd(p.rhs, p.rhs_idx, i);
            const bool lt = x < y;
            const int64_t diff = lt ? y - x : y - x + (int64_t)P;  // prim.rs:1209-1213
            const uint32_t d32 = (uint32_t)diff;
            p.out[i] = lt ? 4096 : 0;
            uint32_t* r = head2(row, p, i);
            *r++ = to_m31(x); *r++ = to_m31(y); *r++ = lt ? 4096u : 0u;
            *r++ = (uint32_t)(((diff % (int64_t)P) + (int64_t)P) % (int64_t)P);
            *r++ = lt ? 0u : 1u;  // borrow
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t limb = (d32 >> (8 * k)) & 0xFFu;
                *r++ = limb;
                atomicAdd(&hist[limb], 1u);
            }
            *r++ = M1; *r++ = M1; *r++ = mult; *r = 1u;
        } else if (OP == LB_OP_RECIP || OP == LB_OP_SQRT) {  // recip/table.rs, sqrt/table.rs: 13 columns
            const int64_t x = rd(p.lhs, p.lhs_idx, i);
            int64_t o, rem;
            if (OP == LB_OP_RECIP) {
                o = x ? floor_div((int64_t)1 << 24, x) : 0;
                rem = ((int64_t)1 << 24) - o * x;
            } else {
                const int64_t t = x << 12;
                o = isqrt_floor(t);
                rem = t - o * o;
            }
            p.out[i] = (int32_t)o;
            uint32_t* r = head1(row, p, i);
            *r++ = to_m31(x); *r++ = to_m31(o); *r++ = to_m31(rem); *r++ = 4096u; *r++ = M1; *r = mult;
        } else if (OP == LB_OP_SIN || OP == LB_OP_EXP2 || OP == LB_OP_LOG2) {  // sin/exp2/log2 table.rs: 12 columns
            const int64_t x = rd(p.lhs, p.lhs_idx, i);
            // LookupLayout::find_index (preprocessed.rs:96-116)
            uint32_t at = 0xFFFFFFFFu;
#pragma unroll 1
            for (int k = 0; k < p.lut.n; ++k)
                if (x >= p.lut.lo[k] && x <= p.lut.hi[k]) at = p.lut.base[k] + (uint32_t)(x - p.lut.lo[k]);
            uint32_t fv = 0;
            if (at == 0xFFFFFFFFu) {
                atomicExch(p.err, 1);  // "Value should fit in range."
            } else {
                fv = p.lut_vals[at];
                atomicAdd(&p.lut_mult[at], 1u);
            }
            p.out[i] = fv > P / 2 ? (int32_t)((int64_t)fv - (int64_t)P) : (int32_t)fv;
            uint32_t* r = head1(row, p, i);
            *r++ = to_m31(x); *r++ = fv; *r++ = M1; *r++ = mult; *r = 1u;
        } else if (OP == LB_OP_CONTIGUOUS) {  // contiguous/table.rs: 11 columns
            const int64_t x = rd(p.lhs, p.lhs_idx, i);
            p.out[i] = (int32_t)x;
            uint32_t* r = head1(row, p, i);
            *r++ = to_m31(x); *r++ = to_m31(x); *r++ = M1; *r = mult;
        }
    }
    __syncthreads();
    if (OP == LB_OP_LESS_THAN) {
        const uint32_t c = hist[threadIdx.x];
        if (c) atomicAdd(&p.lut_mult[threadIdx.x], c);
    }
    // copy-out of the CTA's run of rows
    const uint64_t r0 = blockIdx.x * (uint64_t)blockDim.x;
    const uint32_t have = (uint32_t)min((uint64_t)256, p.n - r0);
    const uint32_t words = have * NC;
    uint32_t* dst = p.rows + r0 * NC;
    const uint32_t nvec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (words >> 2) : 0;
    for (uint32_t v = threadIdx.x; v < nvec; v += 256) {
        uint32_t w = 4 * v, r = w / NC, c = w - r * NC;  // NC is a compile-time constant: no division instruction
        uint4 x;
        x.x = stage[r * NCP + c];
        if (++c == NC) { c = 0; ++r; }
        x.y = stage[r * NCP + c];
        if (++c == NC) { c = 0; ++r; }
        x.z = stage[r * NCP + c];
        if (++c == NC) { c = 0; ++r; }
        x.w = stage[r * NCP + c];
        reinterpret_cast<uint4*>(dst)[v] = x;
    }
    for (uint32_t w = 4 * nvec + threadIdx.x; w < words; w += 256) {
        const uint32_t r = w / NC, c = w - r * NC;
        dst[w] = stage[r * NCP + c];
    }
}

// reductions: output element i folds `group` gathered inputs; one row per step (sum_reduce/table.rs 14 columns,
// max_reduce/table.rs 15 columns).  The finished value appears in every row, hence the two sweeps.
// STAGED (group * NC words per thread fit the 48 KB of static-limit shared memory): the CTA's 256 * group rows are one
// contiguous run, parked in shared memory and stored coalesced like trace_rows_kernel does.
template <bool MAX, bool STAGED>
__global__ void __launch_bounds__(256) trace_reduce_kernel(const TraceOp p) {
    pdl_wait();
    extern __shared__ uint32_t rstage[];
    constexpr int NC = MAX ? 15 : 14;
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t g = p.group, base = i * g;
    const uint32_t pitch = ((uint32_t)g * NC) | 1;
    if (i < p.n) {
        int64_t fin = MAX ? rd(p.lhs, p.lhs_idx, base) : 0;
        for (uint64_t j = 0; j < g; ++j) {
            const int64_t x = rd(p.lhs, p.lhs_idx, base + j);
            fin = MAX ? (x > fin ? x : fin) : fin + x;
        }
        p.out[i] = (int32_t)fin;
        const uint32_t mult = p.out_mult ? p.out_mult[i] : p.out_mult_all;
        int64_t acc = 0;  // running value before the step (0 before the first step, also for max: prim.rs:1703-1709)
        for (uint64_t j = 0; j < g; ++j) {
            const int64_t x = rd(p.lhs, p.lhs_idx, base + j);
            const bool is_max = (j == 0) || x > acc;
            const int64_t next = MAX ? (is_max ? x : acc) : acc + x;
            const bool last_step = (j == g - 1);
            uint32_t* r = head1(STAGED ? rstage + threadIdx.x * pitch + (uint32_t)j * NC : p.rows + (base + j) * NC, p, i);
            *r++ = to_m31(x); *r++ = to_m31(fin); *r++ = to_m31(acc); *r++ = to_m31(next); *r++ = last_step;
            if (MAX) *r++ = is_max;
            *r++ = P - 1; *r = last_step ? mult : 0u;
            acc = next;
        }
    }
    if (!STAGED) return;
    __syncthreads();
    const uint64_t o0 = blockIdx.x * (uint64_t)blockDim.x;
    const uint32_t have = (uint32_t)min((uint64_t)256, p.n - o0);
    const uint32_t per = (uint32_t)g * NC, words = have * per;
    uint32_t* dst = p.rows + o0 * per;
    const uint32_t nvec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (words >> 2) : 0;
    if (nvec) {
        const uint32_t dt = 1024 / per, dc = 1024 - dt * per;
        uint32_t t = (4 * threadIdx.x) / per, c = 4 * threadIdx.x - t * per;
        for (uint32_t v = threadIdx.x; v < nvec; v += 256) {
            uint32_t tt = t, cc = c;
            uint4 x;
            x.x = rstage[tt * pitch + cc];
            if (++cc == per) { cc = 0; ++tt; }
            x.y = rstage[tt * pitch + cc];
            if (++cc == per) { cc = 0; ++tt; }
            x.z = rstage[tt * pitch + cc];
            if (++cc == per) { cc = 0; ++tt; }
            x.w = rstage[tt * pitch + cc];
            reinterpret_cast<uint4*>(dst)[v] = x;
            t += dt;
            c += dc;
            if (c >= per) { c -= per; ++t; }
        }
    }
    for (uint32_t w = 4 * nvec + threadIdx.x; w < words; w += 256) {
        const uint32_t t = w / per, c = w - t * per;
        dst[w] = rstage[t * pitch + c];
    }
}

__global__ void __launch_bounds__(256) count_uses_kernel(uint32_t* __restrict__ uses, const uint32_t* __restrict__ idx, uint64_t n) {
    pdl_wait();
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&uses[idx ? (uint64_t)idx[i] : i], 1u);
}

}  // namespace

cudaError_t trace_count_uses(uint32_t* uses, const uint32_t* idx, uint64_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    launch_k(count_uses_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, uses, idx, n);
    return cudaGetLastError();
}

cudaError_t trace_op(const TraceOp& p, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((p.n + 255) / 256);
#define LB_TRACE_CASE(K) \
    case K: launch_k(trace_rows_kernel<K>, blocks, 256, 0, stream, p); break;
    switch (p.op) {
        LB_TRACE_CASE(LB_OP_ADD) LB_TRACE_CASE(LB_OP_MUL) LB_TRACE_CASE(LB_OP_REM) LB_TRACE_CASE(LB_OP_LESS_THAN)
        LB_TRACE_CASE(LB_OP_RECIP) LB_TRACE_CASE(LB_OP_SQRT) LB_TRACE_CASE(LB_OP_SIN) LB_TRACE_CASE(LB_OP_EXP2)
        LB_TRACE_CASE(LB_OP_LOG2) LB_TRACE_CASE(LB_OP_CONTIGUOUS) LB_TRACE_CASE(LB_OP_INPUTS)
        case LB_OP_SUM_REDUCE:
        case LB_OP_MAX_REDUCE: {
            const bool mx = p.op == LB_OP_MAX_REDUCE;
            const size_t smem = 256 * (size_t)((p.group * (mx ? 15 : 14)) | 1) * sizeof(uint32_t);
            if (p.group <= 3 && smem <= 48 * 1024) {
                if (mx) launch_k(trace_reduce_kernel<true, true>, blocks, 256, smem, stream, p);
                else launch_k(trace_reduce_kernel<false, true>, blocks, 256, smem, stream, p);
            } else {
                if (mx) launch_k(trace_reduce_kernel<true, false>, blocks, 256, 0, stream, p);
                else launch_k(trace_reduce_kernel<false, false>, blocks, 256, 0, stream, p);
            }
            break;
        }
        default: return cudaErrorInvalidValue;
    }
#undef LB_TRACE_CASE
    return cudaGetLastError();
}

}  // namespace lb
