// Streaming kernels of the polynomial-commitment / FRI stages for sm_100a.
//
// Each replaces one stwo SimdBackend trait method reached from stwo::prover::prove, which
// /root/reference calls at crates/prover/src/prover.rs:311-312:
//   eval_at_point          PolyOps::eval_at_point           (OODS sampling)
//   accumulate_quotients   QuotientOps::accumulate_quotients (DEEP quotients, a11)
//   fold_circle_into_line / fold_line   FriOps               (a12)
//   grind_range            GrindOps<Blake2sChannel>::grind   (a13)
// All of them are one-pass, HBM-bound kernels over column-major u32 data: a warp reads 128
// contiguous bytes per column, every input byte is read once, no shared-memory staging is
// needed except for the block reductions of eval_at_point.
#include "launch.cuh"
#include "kernels.cuh"

#include "blake2s.cuh"

namespace lb {

__constant__ Pt k_gen_pow2[31];  // G * 2^j, G = (2, 1268011823)

cudaError_t kernels_init(cudaStream_t stream) {
    Pt hp[31];
    Pt p = {2, 1268011823u};
    for (int j = 0; j < 31; ++j) {
        hp[j] = p;
        p = pt_add(p, p);
    }
    cudaError_t e = cudaMemcpyToSymbolAsync(k_gen_pow2, hp, sizeof(hp), 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);  // hp is a stack buffer
}

__device__ __forceinline__ Pt k_point_of_index(uint32_t idx) {
    Pt r = {1, 0};
#pragma unroll 1
    for (int j = 0; j < 31; ++j)
        if ((idx >> j) & 1) r = pt_add(r, k_gen_pow2[j]);
    return r;
}

// ------------------------------------------------------------------------------------
// eval_at_point:  f(P) = sum_k c_k * prod_{bit b of k} mappings[b]
// The index is split k = (chunk << m) | e.  basis[e] covers the low m bits and is shared by all
// columns and chunks; each CTA reduces one chunk of one column against it, multiplies by the
// chunk's high-bit factor and stores one QM31 partial; a second kernel sums the partials.
// Traffic: 4 B per coefficient (+ the 64 KiB basis table, L1/L2 resident).
// ------------------------------------------------------------------------------------
__global__ void eval_basis_kernel(QM31* basis, const QM31* mappings, int m) {
    pdl_wait();
    pdl_launch_dependents();
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (1u << m)) return;
    QM31 r = q_from_m(1);
    for (int b = 0; b < m; ++b)
        if ((k >> b) & 1) r = q_mul(r, mappings[b]);
    basis[k] = r;
}

__device__ __forceinline__ uint32_t fold64(uint64_t x) {
    // any u64 -> canonical M31
    uint64_t y = (x & P) + (x >> 31);  // < 2^34
    return m_reduce64(y);
}

__device__ __forceinline__ QM31 warp_sum_q(QM31 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        QM31 w;
        w.a.a = __shfl_xor_sync(0xFFFFFFFFu, v.a.a, o);
        w.a.b = __shfl_xor_sync(0xFFFFFFFFu, v.a.b, o);
        w.b.a = __shfl_xor_sync(0xFFFFFFFFu, v.b.a, o);
        w.b.b = __shfl_xor_sync(0xFFFFFFFFu, v.b.b, o);
        v = q_add(v, w);
    }
    return v;
}

// block-wide QM31 sum (blockDim.x == 256); result valid in thread 0
__device__ __forceinline__ QM31 block_sum_q(QM31 v) {
    __shared__ QM31 s_w[8];
    v = warp_sum_q(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_w[w] = v;
    __syncthreads();
    QM31 r = q_zero();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r = q_add(r, s_w[i]);
    }
    __syncthreads();
    return r;
}

// One CTA = one chunk of 2^m coefficients x up to EVAL_CPB columns.  The chunk's basis entries (16 per thread, 4 runs of 4
// consecutive ones) are loaded ONCE into registers and reused for every column of the group: the first version re-read the
// 16-byte basis entry for every 4-byte coefficient, i.e. 4x the column traffic out of L2 (ncu r2: 17 % of DRAM peak, 13.8
// long-scoreboard stalls per issue).  Coefficients come in as 128-bit loads, four in flight per thread and column.
constexpr int EVAL_CPB = 8;
__global__ void __launch_bounds__(256) eval_partial_kernel(const uint32_t* const* __restrict__ cols, int n_cols, int log, int m,
                                                           const QM31* __restrict__ basis,
                                                           const QM31* __restrict__ mappings, QM31* __restrict__ partials) {
    pdl_wait();
    const uint32_t chunk = blockIdx.x;
    const uint32_t n_chunks = gridDim.x;
    const int c0 = blockIdx.y * EVAL_CPB;
    const int nc = min(EVAL_CPB, n_cols - c0);
    const uint32_t elems = 1u << m;  // 2^EVAL_CHUNK_LOG on this path
    const uint32_t tid = threadIdx.x;
    __shared__ uint4 s_part[EVAL_CPB][256];  // 32 KiB
    // thread t owns elements 4 * (t + 256 i) + q, i < NI, q < 4
    constexpr int NI = (1 << EVAL_CHUNK_LOG) / 1024;
    uint4 bs[4 * NI];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) bs[4 * i + q] = *reinterpret_cast<const uint4*>(&basis[4 * (tid + 256 * i) + q]);
    // the chunk's high-bit factor (the same for every column)
    QM31 hi = q_from_m(1);
    for (int b = 0; m + b < log; ++b)
        if ((chunk >> b) & 1) hi = q_mul(hi, mappings[m + b]);
    uint4 v[NI], nv[NI];
    {
        const uint4* col = reinterpret_cast<const uint4*>(cols[c0] + ((size_t)chunk << m));
#pragma unroll
        for (int i = 0; i < NI; ++i) v[i] = col[tid + 256 * i];
    }
    for (int c = 0; c < nc; ++c) {
        if (c + 1 < nc) {  // the next column's loads are in flight while this one is multiplied
            const uint4* col = reinterpret_cast<const uint4*>(cols[c0 + c + 1] + ((size_t)chunk << m));
#pragma unroll
            for (int i = 0; i < NI; ++i) nv[i] = col[tid + 256 * i];
        }
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint64_t cq = w[q];
                const uint4 b = bs[4 * i + q];
                a0 += cq * b.x;
                a1 += cq * b.y;
                a2 += cq * b.z;
                a3 += cq * b.w;
                if (q == 3) {  // canonical operands: four products (< 2^64 - 2^34) plus a folded sum (< 2^33 + 2^31) fit 64 bits
                    a0 = (a0 & P) + (a0 >> 31);
                    a1 = (a1 & P) + (a1 >> 31);
                    a2 = (a2 & P) + (a2 >> 31);
                    a3 = (a3 & P) + (a3 >> 31);
                }
            }
        }
        // every thread parks its sum in shared memory; the block-wide reduction happens once, after the last column (no barrier
        // and no shuffle tree per column: the next column's loads are not held back by this one's reduction)
        s_part[c][tid] = make_uint4(fold64(a0), fold64(a1), fold64(a2), fold64(a3));
#pragma unroll
        for (int i = 0; i < NI; ++i) v[i] = nv[i];
    }
    __syncthreads();
    {
        // warp w reduces column w: 8 entries per lane, then one shuffle tree
        const int w = tid >> 5, lane = tid & 31;
        if (w < nc) {
            QM31 r = q_zero();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 v = s_part[w][lane + 32 * k];
                r = q_add(r, q_make(v.x, v.y, v.z, v.w));
            }
            r = warp_sum_q(r);
            if (lane == 0) partials[(size_t)(c0 + w) * n_chunks + chunk] = q_mul(r, hi);
        }
    }
    (void)elems;
}

// columns of fewer than 2^EVAL_CHUNK_LOG coefficients: one CTA per column, scalar loads
__global__ void __launch_bounds__(256) eval_partial_small_kernel(const uint32_t* const* __restrict__ cols, int log, int m,
                                                                 const QM31* __restrict__ basis, QM31* __restrict__ partials) {
    pdl_wait();
    const uint32_t* col = cols[blockIdx.y];
    const uint32_t elems = 1u << m;
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int pending = 0;
    for (uint32_t e = threadIdx.x; e < elems; e += 256) {
        uint64_t c = col[e];
        uint4 b = *reinterpret_cast<const uint4*>(&basis[e]);
        a0 += c * b.x;
        a1 += c * b.y;
        a2 += c * b.z;
        a3 += c * b.w;
        if (++pending == 4) {  // see eval_partial_kernel
            a0 = (a0 & P) + (a0 >> 31);
            a1 = (a1 & P) + (a1 >> 31);
            a2 = (a2 & P) + (a2 >> 31);
            a3 = (a3 & P) + (a3 >> 31);
            pending = 0;
        }
    }
    QM31 v = q_make(fold64(a0), fold64(a1), fold64(a2), fold64(a3));
    v = block_sum_q(v);
    if (threadIdx.x == 0) partials[blockIdx.y] = v;
    (void)log;
}

__global__ void __launch_bounds__(256) eval_sum_kernel(const QM31* __restrict__ partials, uint32_t n_chunks, QM31* out) {
    pdl_wait();
    pdl_launch_dependents();
    const QM31* p = partials + (size_t)blockIdx.x * n_chunks;
    QM31 v = q_zero();
    for (uint32_t i = threadIdx.x; i < n_chunks; i += 256) v = q_add(v, p[i]);
    v = block_sum_q(v);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
}

cudaError_t eval_at_point(const uint32_t* const* d_cols, int n_cols, int log, const QM31* d_mappings, QM31* d_basis,
                          QM31* d_partials, QM31* d_out, cudaStream_t stream) {
    if (n_cols == 0) return cudaSuccess;
    int m = log < EVAL_CHUNK_LOG ? log : EVAL_CHUNK_LOG;
    uint32_t n_chunks = 1u << (log - m);
    uint32_t nb = 1u << m;
    launch_k(eval_basis_kernel, (nb + 255) / 256, 256, 0, stream, d_basis, d_mappings, m);
    if (m == EVAL_CHUNK_LOG) {
        dim3 grid(n_chunks, (n_cols + EVAL_CPB - 1) / EVAL_CPB);
        launch_k(eval_partial_kernel, grid, 256, 0, stream, d_cols, n_cols, log, m, d_basis, d_mappings, d_partials);
    } else {
        dim3 grid(1, n_cols);
        launch_k(eval_partial_small_kernel, grid, 256, 0, stream, d_cols, log, m, d_basis, d_partials);
    }
    launch_k(eval_sum_kernel, n_cols, 256, 0, stream, d_partials, n_chunks, d_out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// DEEP quotients.  Per row:  acc = sum over batches (Horner in rc_pow) of
//     ( sum_j c_j * f_j(row) - (A * y + B) ) * 1 / ((Pr.x - x) Pi.y - (Pr.y - y) Pi.x)
// Every column value is read once per batch it belongs to (4 B per column per row + 16 B written).
// ------------------------------------------------------------------------------------
// Domain points come from the CFFT twiddle tables: storage row 2h is (x_h, y_h) with y_h = Y[log-1][h] and
// x_h = +-X[log-1][h >> 1] (sign by the low bit of h: adjacent half-coset points differ by the order-2 point),
// row 2h+1 is the conjugate.  Each thread handles QROWS rows 256 apart so that all QROWS * n_batches
// denominators share ONE field inversion.
#ifndef LB_QROWS
#define LB_QROWS 4
#endif
#ifndef LB_Q_MINBLOCKS
#define LB_Q_MINBLOCKS 3  // measured r2 (cfg 3 DEEP stage): 2 -> 0.670 ms, 3 -> 0.632 ms, 4 rows x 3 CTAs; 2 rows x 4 CTAs 0.645, 8 rows x 1 CTA 0.791
#endif
constexpr int QROWS = LB_QROWS;

template <int NB>
__global__ void __launch_bounds__(256, LB_Q_MINBLOCKS) quotients_kernel(uint32_t* __restrict__ o0, uint32_t* __restrict__ o1,
                                                        uint32_t* __restrict__ o2, uint32_t* __restrict__ o3,
                                                        const uint32_t* const* __restrict__ cols,
                                                        const QuotientEntry* __restrict__ entries,
                                                        const __grid_constant__ QuotientParams qp,
                                                        const uint2* __restrict__ tw_x, const uint2* __restrict__ tw_y,
                                                        uint32_t n, uint32_t row0) {
    pdl_wait();
    // rows [row0, row0 + n) of the domain: the columns and the outputs are this row range (the whole domain on one GPU;
    // a rank's row shard in the sharded prover), the domain points are those of the global rows
    const uint32_t j0 = blockIdx.x * (256 * QROWS) + threadIdx.x;
    uint32_t ys[QROWS];
    CM31 den[QROWS * NB], pre[QROWS * NB];
    CM31 run = {1, 0};
#pragma unroll
    for (int r = 0; r < QROWS; ++r) {
        uint32_t j = j0 + r * 256;
        if (j >= n) j = n - 1;  // clamp (results of clamped rows are not stored)
        j += row0;
        uint32_t h = j >> 1;
        uint32_t x = tw_x[h >> 1].x;
        if (h & 1) x = m_neg(x);
        uint32_t y = tw_y[h].x;
        if (j & 1) y = m_neg(y);
        ys[r] = y;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const QuotientBatch& B = qp.b[b];
            CM31 dx = {m_sub(B.prx.a, x), B.prx.b};
            CM31 dy = {m_sub(B.pry.a, y), B.pry.b};
            CM31 d = c_sub(c_mul(dx, B.piy), c_mul(dy, B.pix));
            den[r * NB + b] = d;
            pre[r * NB + b] = run;
            run = c_mul(run, d);
        }
    }
    CM31 inv = c_inv(run);
#pragma unroll
    for (int k = QROWS * NB - 1; k >= 0; --k) {
        CM31 di = c_mul(inv, pre[k]);
        inv = c_mul(inv, den[k]);
        den[k] = di;  // now the inverse
    }
    // Numerators sum_j c_j * f_j(row): the four coordinates of every row are kept as lazy 64-bit sums (one IMAD.WIDE per
    // coordinate and column; products are < 2^62, so a fold every third column keeps them below 2^64) and reduced once per
    // batch.  The QROWS rows of a thread go through the column loop together: one entry / pointer load serves all of them
    // and the rows give the loads and multiply-adds independent chains.
    uint32_t jr[QROWS];
#pragma unroll
    for (int r = 0; r < QROWS; ++r) {
        uint32_t j = j0 + r * 256;
        jr[r] = j < n ? j : n - 1;  // clamp (results of clamped rows are not stored)
    }
    QM31 acc[QROWS];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const QuotientBatch& B = qp.b[b];
        uint64_t s[QROWS][4];
#pragma unroll
        for (int r = 0; r < QROWS; ++r) s[r][0] = s[r][1] = s[r][2] = s[r][3] = 0;
        // columns go three at a time: 3 x QROWS independent loads in flight per thread (the kernel is latency-bound
        // otherwise: ~100 registers per thread leave 16 warps per SM), then one fold of the twelve products
        int k = 0;
        for (; k + 3 <= B.count; k += 3) {
            const QuotientEntry e0 = entries[B.first + k], e1 = entries[B.first + k + 1], e2 = entries[B.first + k + 2];
            const uint32_t* __restrict__ c0 = cols[e0.col];
            const uint32_t* __restrict__ c1 = cols[e1.col];
            const uint32_t* __restrict__ c2 = cols[e2.col];
            uint32_t v0[QROWS], v1[QROWS], v2[QROWS];
#pragma unroll
            for (int r = 0; r < QROWS; ++r) {
                v0[r] = c0[jr[r]];
                v1[r] = c1[jr[r]];
                v2[r] = c2[jr[r]];
            }
#pragma unroll
            for (int r = 0; r < QROWS; ++r) {
                const uint64_t a = v0[r], b2 = v1[r], c = v2[r];
                s[r][0] += a * e0.c.a.a + b2 * e1.c.a.a + c * e2.c.a.a;
                s[r][1] += a * e0.c.a.b + b2 * e1.c.a.b + c * e2.c.a.b;
                s[r][2] += a * e0.c.b.a + b2 * e1.c.b.a + c * e2.c.b.a;
                s[r][3] += a * e0.c.b.b + b2 * e1.c.b.b + c * e2.c.b.b;
#pragma unroll
                for (int q = 0; q < 4; ++q) s[r][q] = (s[r][q] & P) + (s[r][q] >> 31);
            }
        }
        for (; k < B.count; ++k) {  // at most two more columns: folded sums (< 2^34) + 2 products (< 2^63) fit
            const QuotientEntry e = entries[B.first + k];
            const uint32_t* __restrict__ col = cols[e.col];
#pragma unroll
            for (int r = 0; r < QROWS; ++r) {
                const uint64_t v = col[jr[r]];
                s[r][0] += v * e.c.a.a;
                s[r][1] += v * e.c.a.b;
                s[r][2] += v * e.c.b.a;
                s[r][3] += v * e.c.b.b;
            }
        }
#pragma unroll
        for (int r = 0; r < QROWS; ++r) {
            QM31 num = q_make(fold64(s[r][0]), fold64(s[r][1]), fold64(s[r][2]), fold64(s[r][3]));
            QM31 lin = q_add(q_mul_m(B.sum_a, ys[r]), B.sum_b);
            num = q_sub(num, lin);
            QM31 q = q_mul_c(num, den[r * NB + b]);
            acc[r] = (b == 0) ? q : q_add(q_mul(acc[r], B.rc_pow), q);
        }
    }
#pragma unroll
    for (int r = 0; r < QROWS; ++r) {
        uint32_t j = j0 + r * 256;
        if (j >= n) continue;
        o0[j] = acc[r].a.a;
        o1[j] = acc[r].a.b;
        o2[j] = acc[r].b.a;
        o3[j] = acc[r].b.b;
    }
}

cudaError_t accumulate_quotients(uint32_t* const out[4], const uint32_t* const* d_cols, const QuotientEntry* d_entries,
                                 const QuotientParams& qp, const Twiddles* tw, int log, cudaStream_t stream, uint32_t row0,
                                 uint32_t n_rows) {
    if (qp.n_batches < 1 || qp.n_batches > MAX_QUOTIENT_BATCHES) return cudaErrorInvalidValue;
    if (log < 2 || log > tw->max_log) return cudaErrorInvalidValue;
    uint32_t n = n_rows ? n_rows : (1u << log);
    if ((uint64_t)row0 + n > ((uint64_t)1 << log) || (row0 & 1) || (n & 1)) return cudaErrorInvalidValue;
    const uint2* tw_x = tw->fwd + ((size_t)1 << (log - 2));              // X[log-1]
    const uint2* tw_y = tw->fwd + tw->y_off + ((size_t)1 << (log - 1));  // Y[log-1]
    unsigned blocks = (n + 256 * QROWS - 1) / (256 * QROWS);
#define LB_Q_LAUNCH(NB) \
    launch_k(quotients_kernel<NB>, blocks, 256, 0, stream, out[0], out[1], out[2], out[3], d_cols, d_entries, qp, tw_x, tw_y, n, row0)
    switch (qp.n_batches) {
        case 1: LB_Q_LAUNCH(1); break;
        case 2: LB_Q_LAUNCH(2); break;
        case 3: LB_Q_LAUNCH(3); break;
        default: LB_Q_LAUNCH(4); break;
    }
#undef LB_Q_LAUNCH
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// FRI folds (48 B per output element: 32 read + 16 written, + dst read for the circle fold)
// ------------------------------------------------------------------------------------
struct Coords4 {
    uint32_t* p[4];
};
struct CCoords4 {
    const uint32_t* p[4];
};

template <bool CIRCLE>
__global__ void __launch_bounds__(256) fold_kernel(Coords4 dst, CCoords4 src, const uint2* __restrict__ itw, uint32_t n_out,
                                                   QM31 alpha, QM31 alpha_sq) {
    pdl_wait();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    uint32_t a[4], b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint2 v = *reinterpret_cast<const uint2*>(src.p[c] + 2 * (size_t)i);
        a[c] = v.x;
        b[c] = v.y;
    }
    uint32_t tinv = itw[i].x;
    QM31 f_p = q_make(a[0], a[1], a[2], a[3]), f_n = q_make(b[0], b[1], b[2], b[3]);
    QM31 f0 = q_add(f_p, f_n);
    QM31 f1 = q_mul_m(q_sub(f_p, f_n), tinv);
    QM31 r = q_add(q_mul(alpha, f1), f0);
    if (CIRCLE) {
        QM31 d = q_make(dst.p[0][i], dst.p[1][i], dst.p[2][i], dst.p[3][i]);
        r = q_add(q_mul(d, alpha_sq), r);
    }
    dst.p[0][i] = r.a.a;
    dst.p[1][i] = r.a.b;
    dst.p[2][i] = r.b.a;
    dst.p[3][i] = r.b.b;
}

cudaError_t fold_circle_into_line(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log,
                                  QM31 alpha, cudaStream_t stream) {
    Coords4 d;
    CCoords4 s;
    for (int c = 0; c < 4; ++c) {
        d.p[c] = dst[c];
        s.p[c] = src[c];
    }
    uint32_t n_out = 1u << (log - 1);
    launch_k(fold_kernel<true>, (n_out + 255) / 256, 256, 0, stream, d, s, itw, n_out, alpha, q_mul(alpha, alpha));
    return cudaGetLastError();
}

cudaError_t fold_line(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log, QM31 alpha,
                      cudaStream_t stream) {
    Coords4 d;
    CCoords4 s;
    for (int c = 0; c < 4; ++c) {
        d.p[c] = dst[c];
        s.p[c] = src[c];
    }
    uint32_t n_out = 1u << (log - 1);
    launch_k(fold_kernel<false>, (n_out + 255) / 256, 256, 0, stream, d, s, itw, n_out, alpha, q_zero());
    return cudaGetLastError();
}

// the same folds with the folding coefficient read from device memory (drawn there by the device-side channel)
template <bool CIRCLE>
__global__ void __launch_bounds__(256) fold_kernel_dev_alpha(Coords4 dst, CCoords4 src, const uint2* __restrict__ itw,
                                                             uint32_t n_out, const QM31* __restrict__ alpha_ptr) {
    pdl_wait();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const QM31 alpha = *alpha_ptr;
    uint32_t a[4], b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint2 v = *reinterpret_cast<const uint2*>(src.p[c] + 2 * (size_t)i);
        a[c] = v.x;
        b[c] = v.y;
    }
    uint32_t tinv = itw[i].x;
    QM31 f_p = q_make(a[0], a[1], a[2], a[3]), f_n = q_make(b[0], b[1], b[2], b[3]);
    QM31 f0 = q_add(f_p, f_n);
    QM31 f1 = q_mul_m(q_sub(f_p, f_n), tinv);
    QM31 r = q_add(q_mul(alpha, f1), f0);
    if (CIRCLE) {
        QM31 d = q_make(dst.p[0][i], dst.p[1][i], dst.p[2][i], dst.p[3][i]);
        r = q_add(q_mul(d, q_mul(alpha, alpha)), r);
    }
    dst.p[0][i] = r.a.a;
    dst.p[1][i] = r.a.b;
    dst.p[2][i] = r.b.a;
    dst.p[3][i] = r.b.b;
}

cudaError_t fold_circle_into_line_dev(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log,
                                      const QM31* d_alpha, cudaStream_t stream) {
    Coords4 d;
    CCoords4 s;
    for (int c = 0; c < 4; ++c) {
        d.p[c] = dst[c];
        s.p[c] = src[c];
    }
    uint32_t n_out = 1u << (log - 1);
    launch_k(fold_kernel_dev_alpha<true>, (n_out + 255) / 256, 256, 0, stream, d, s, itw, n_out, d_alpha);
    return cudaGetLastError();
}

cudaError_t fold_line_dev(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log, const QM31* d_alpha,
                          cudaStream_t stream) {
    Coords4 d;
    CCoords4 s;
    for (int c = 0; c < 4; ++c) {
        d.p[c] = dst[c];
        s.p[c] = src[c];
    }
    uint32_t n_out = 1u << (log - 1);
    launch_k(fold_kernel_dev_alpha<false>, (n_out + 255) / 256, 256, 0, stream, d, s, itw, n_out, d_alpha);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// Device-side Fiat-Shamir step of the FRI commit loop (FriProver::commit: per inner layer
// `mix_root(layer root); alpha = draw_secure_felt()`), so that fold -> Merkle -> mix -> draw -> fold runs
// on the stream without a host round trip per layer.  Same Blake2sChannel as prover.cu's host class
// (core/channel/blake2s.rs; variant 0 "legacy" / 1 "v2").  One thread.
// ------------------------------------------------------------------------------------
__device__ void channel_mix_root_draw_dev(DevChannel* ch, const uint32_t* root, int variant, QM31* alpha_out,
                                         uint32_t* digest_log) {
    uint32_t h[8], m[16];
    // mix_root: digest <- Blake2s(digest || root), one 64-byte final block
    blake2s_init(h);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m[i] = ch->digest[i];
        m[8 + i] = root[i];
    }
    blake2s_compress(h, m, 64, 0, 0xFFFFFFFFu);
    uint32_t dg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dg[i] = h[i];
        ch->digest[i] = h[i];
        digest_log[i] = h[i];
    }
    // draw_secure_felt: draw_base_felts retries until all eight words are < 2p
    uint32_t n_sent = 0;
    for (;;) {
        blake2s_init(h);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = dg[i];
        m[8] = n_sent;
#pragma unroll
        for (int i = 9; i < 16; ++i) m[i] = 0;
        ++n_sent;
        if (variant == 0) {
            blake2s_compress(h, m, 64, 0, 0xFFFFFFFFu);
        } else {
            blake2s_compress(h, m, 64, 0, 0);  // 65-byte message: a full block, then the 0x00 domain byte
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = 0;
            blake2s_compress(h, m, 65, 0, 0xFFFFFFFFu);
        }
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; ++i) ok = ok && (h[i] < 2 * P);
        if (ok) break;
    }
    ch->n_sent = n_sent;
    *alpha_out = q_make(h[0] >= P ? h[0] - P : h[0], h[1] >= P ? h[1] - P : h[1], h[2] >= P ? h[2] - P : h[2],
                        h[3] >= P ? h[3] - P : h[3]);
}

__global__ void channel_mix_root_draw_kernel(DevChannel* ch, const uint32_t* root, int variant, QM31* alpha_out,
                                             uint32_t* digest_log) {
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    channel_mix_root_draw_dev(ch, root, variant, alpha_out, digest_log);
}

// Row-sharded layer (sharded prover): `roots` holds the all-gathered sub-tree roots of the W = 2^logw ranks; the top logw
// levels of the layer tree are hashed here (node = Blake2s(left || right)), then mix_root + draw as above.  `top_out`
// receives every level: level k (2^k digests) at word 8 * (2^k - 1), level logw = the gathered roots themselves.
__global__ void channel_mix_sharded_root_draw_kernel(DevChannel* ch, const uint32_t* roots, int logw, int variant,
                                                     QM31* alpha_out, uint32_t* digest_log, uint32_t* top_out) {
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int w = 1 << logw;
    for (int i = 0; i < 8 * w; ++i) top_out[8 * (w - 1) + i] = roots[i];
    for (int k = logw - 1; k >= 0; --k) {
        const uint32_t* child = top_out + 8 * ((2 << k) - 1);
        uint32_t* lvl = top_out + 8 * ((1 << k) - 1);
        for (int i = 0; i < (1 << k); ++i) {
            uint32_t h[8], m[16];
            blake2s_init(h);
            for (int q = 0; q < 16; ++q) m[q] = child[16 * i + q];
            blake2s_compress(h, m, 64, 0, 0xFFFFFFFFu);
            for (int q = 0; q < 8; ++q) lvl[8 * i + q] = h[q];
        }
    }
    channel_mix_root_draw_dev(ch, top_out, variant, alpha_out, digest_log);
}
cudaError_t channel_mix_sharded_root_draw(DevChannel* d_ch, const uint32_t* d_roots, int logw, int variant, QM31* d_alpha_out,
                                          uint32_t* d_digest_log, uint32_t* d_top_out, cudaStream_t stream) {
    if (logw < 1 || logw > 6) return cudaErrorInvalidValue;
    launch_k(channel_mix_sharded_root_draw_kernel, 1, 32, 0, stream, d_ch, d_roots, logw, variant, d_alpha_out, d_digest_log, d_top_out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// The tail of the FRI commit loop in one launch.  Once a line layer has at most 2^FRI_TAIL_MAX_LOG values, every step of
// FriProver::commit for it - Merkle tree of the layer's four coordinate columns, mix_root, draw the folding coefficient,
// fold_line into the next layer - is a handful of dependent Blake2s compressions and a few hundred field operations: four
// launches of 3-20 us per layer whose cost is launch and drain latency.  One CTA runs all of them for all remaining layers
// with block barriers in between; every tree level and every folded layer is written to the same buffers the layer-by-layer
// path uses (the decommitment reads them later).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) fri_tail_kernel(const __grid_constant__ FriTailArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ QM31 s_alpha;
    __shared__ uint4 s_dig[2 * 1024];  // digest i of the tree level just hashed: s_dig[2 * i], s_dig[2 * i + 1]
    const uint32_t tid = threadIdx.x;
    int step = 0;
    for (int lg = a.from_log; lg > a.last_log; --lg, ++step) {
        const uint32_t n = 1u << lg;
        uint32_t* const T = a.tree[lg];  // layer k of this tree: 2^k digests at word offset 8 * (2^k - 1)
        uint32_t* const V = a.vals[lg];  // coordinate c of this FRI layer at V + c * n
        uint32_t h[8], m[16];
        if (tid < n) {  // leaves: Blake2s-256 of the four coordinate words (one final 16-byte block)
            blake2s_init(h);
#pragma unroll
            for (int c = 0; c < 4; ++c) m[c] = __ldcg(V + (size_t)c * n + tid);
#pragma unroll
            for (int j = 4; j < 16; ++j) m[j] = 0;
            blake2s_compress(h, m, 16, 0, 0xFFFFFFFFu);
            const uint4 lo = make_uint4(h[0], h[1], h[2], h[3]), hi = make_uint4(h[4], h[5], h[6], h[7]);
            uint4* o = reinterpret_cast<uint4*>(T + 8 * ((size_t)n - 1) + 8 * (size_t)tid);
            o[0] = lo;
            o[1] = hi;
            s_dig[2 * tid] = lo;
            s_dig[2 * tid + 1] = hi;
        }
        __syncthreads();
        for (int k = lg - 1; k >= 0; --k) {  // node = Blake2s-256(left || right); children handed over in shared memory
            const bool on = tid < (1u << k);
            uint4 x0, x1, x2, x3;
            if (on) {
                x0 = s_dig[4 * tid]; x1 = s_dig[4 * tid + 1]; x2 = s_dig[4 * tid + 2]; x3 = s_dig[4 * tid + 3];
            }
            __syncthreads();  // all children read: the slots can be overwritten
            if (on) {
                m[0] = x0.x; m[1] = x0.y; m[2] = x0.z; m[3] = x0.w;
                m[4] = x1.x; m[5] = x1.y; m[6] = x1.z; m[7] = x1.w;
                m[8] = x2.x; m[9] = x2.y; m[10] = x2.z; m[11] = x2.w;
                m[12] = x3.x; m[13] = x3.y; m[14] = x3.z; m[15] = x3.w;
                blake2s_init(h);
                blake2s_compress(h, m, 64, 0, 0xFFFFFFFFu);
                const uint4 lo = make_uint4(h[0], h[1], h[2], h[3]), hi = make_uint4(h[4], h[5], h[6], h[7]);
                uint4* o = reinterpret_cast<uint4*>(T + 8 * (((size_t)1 << k) - 1) + 8 * (size_t)tid);
                o[0] = lo;
                o[1] = hi;
                s_dig[2 * tid] = lo;
                s_dig[2 * tid + 1] = hi;
            }
            __syncthreads();
        }
        if (tid == 0) {  // the root was written by this thread
            channel_mix_root_draw_dev(a.ch, T, a.variant, a.alphas + step + 1, a.digests + 8 * (size_t)step);
            s_alpha = a.alphas[step + 1];
        }
        __syncthreads();
        const uint32_t n_out = n >> 1;
        if (tid < n_out) {  // FriOps::fold_line
            const QM31 alpha = s_alpha;
            uint32_t p[4], q[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint2 v = __ldcg(reinterpret_cast<const uint2*>(V + (size_t)c * n + 2 * (size_t)tid));
                p[c] = v.x;
                q[c] = v.y;
            }
            const uint32_t tinv = a.itw[lg][tid].x;
            const QM31 f_p = q_make(p[0], p[1], p[2], p[3]), f_n = q_make(q[0], q[1], q[2], q[3]);
            const QM31 r = q_add(q_mul(alpha, q_mul_m(q_sub(f_p, f_n), tinv)), q_add(f_p, f_n));
            uint32_t* const W = a.vals[lg - 1];
            W[tid] = r.a.a;
            W[(size_t)n_out + tid] = r.a.b;
            W[2 * (size_t)n_out + tid] = r.b.a;
            W[3 * (size_t)n_out + tid] = r.b.b;
        }
        __syncthreads();
    }
}

cudaError_t fri_tail(const FriTailArgs& a, cudaStream_t stream) {
    if (a.from_log > FRI_TAIL_MAX_LOG || a.last_log < 0 || a.from_log <= a.last_log) return cudaErrorInvalidValue;
    launch_k(fri_tail_kernel, 1, 1024, 0, stream, a);
    return cudaGetLastError();
}

cudaError_t channel_mix_root_draw(DevChannel* d_ch, const uint32_t* d_root, int variant, QM31* d_alpha_out, uint32_t* d_digest_log,
                                  cudaStream_t stream) {
    launch_k(channel_mix_root_draw_kernel, 1, 32, 0, stream, d_ch, d_root, variant, d_alpha_out, d_digest_log);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// Proof of work
// ------------------------------------------------------------------------------------
struct Digest8 {
    uint32_t w[8];
};

__global__ void __launch_bounds__(256) grind_kernel(Digest8 dg, int variant, uint32_t pow_bits, uint64_t base,
                                                    uint64_t count, unsigned long long* found) {
    pdl_wait();
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint64_t nonce = base + t;
    uint32_t h[8], m[16];
    uint32_t w0, w1;
    if (variant == 0) {
        // legacy mix_u64: raw compress(h = digest, m = [lo, hi, 0..], t = 0, f = 0)
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = dg.w[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = 0;
        m[0] = (uint32_t)nonce;
        m[1] = (uint32_t)(nonce >> 32);
        blake2s_compress(h, m, 0, 0, 0);
    } else {
        // Blake2s(digest || nonce LE): one 40-byte final block
        blake2s_init(h);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = dg.w[i];
        m[8] = (uint32_t)nonce;
        m[9] = (uint32_t)(nonce >> 32);
#pragma unroll
        for (int i = 10; i < 16; ++i) m[i] = 0;
        blake2s_compress(h, m, 40, 0, 0xFFFFFFFFu);
    }
    w0 = h[0];
    w1 = h[1];
    bool ok;
    if (pow_bits <= 32)
        ok = pow_bits == 32 ? (w0 == 0) : ((w0 & ((1u << pow_bits) - 1)) == 0);
    else
        ok = (w0 == 0) && ((w1 & ((pow_bits >= 64) ? 0xFFFFFFFFu : ((1u << (pow_bits - 32)) - 1))) == 0);
    if (ok) atomicMin(found, (unsigned long long)nonce);
}

cudaError_t grind_range(const uint32_t digest[8], int variant, uint32_t pow_bits, uint64_t base, uint64_t count,
                        unsigned long long* d_found, cudaStream_t stream) {
    Digest8 dg;
    for (int i = 0; i < 8; ++i) dg.w[i] = digest[i];
    uint64_t blocks = (count + 255) / 256;
    launch_k(grind_kernel, (unsigned)blocks, 256, 0, stream, dg, variant, pow_bits, base, count, d_found);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// column utilities
// ------------------------------------------------------------------------------------
__global__ void add_inplace_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t n) {
    pdl_wait();
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = m_add(dst[i], src[i]);
}
cudaError_t add_inplace(uint32_t* dst, const uint32_t* src, size_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    launch_k(add_inplace_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, dst, src, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// Column utilities behind the remaining backend-trait methods (ColumnOps::bit_reverse_column, PolyOps::new_canonical_ordered,
// FieldOps::batch_inverse for BaseField and SecureField columns).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bit_reverse_kernel(uint32_t* __restrict__ col, int log) {
    pdl_wait();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log) return;
    const uint32_t j = log ? __brev(i) >> (32 - log) : 0;
    if (i < j) {
        const uint32_t a = col[i], b = col[j];
        col[i] = b;
        col[j] = a;
    }
}
cudaError_t bit_reverse(uint32_t* col, int log, cudaStream_t stream) {
    const uint32_t n = 1u << log;
    launch_k(bit_reverse_kernel, (n + 255) / 256, 256, 0, stream, col, log);
    return cudaGetLastError();
}

// values in canonic-coset order -> the bit-reversed circle-domain order every other entry point uses
// (CircleEvaluation::new_canonical_ordered: coset point k is circle-domain point k/2 for even k, n/2 + (n-1-k)/2 for odd k)
__global__ void __launch_bounds__(256) canonical_to_storage_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in,
                                                                   int log) {
    pdl_wait();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >> log) return;
    const uint32_t n = 1u << log;
    const uint32_t dom = (k & 1) ? (n >> 1) + ((n - 1 - k) >> 1) : (k >> 1);
    out[log ? __brev(dom) >> (32 - log) : 0] = in[k];
}
cudaError_t canonical_to_storage(uint32_t* out, const uint32_t* in, int log, cudaStream_t stream) {
    const uint32_t n = 1u << log;
    launch_k(canonical_to_storage_kernel, (n + 255) / 256, 256, 0, stream, out, in, log);
    return cudaGetLastError();
}

// four consecutive elements per thread share one field inversion (Montgomery's trick); a zero input raises *flag
__global__ void __launch_bounds__(256) batch_inverse_m31_kernel(uint32_t* out, const uint32_t* in, size_t n, int* flag) {
    pdl_wait();
    const size_t i0 = 4 * (blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    if (i0 >= n) return;
    uint32_t x[4], pre[4];
    uint32_t acc = 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[k] = (i0 + k < n) ? in[i0 + k] : 1u;
        if (x[k] == 0) {
            atomicExch(flag, 1);
            x[k] = 1;
        }
        pre[k] = acc;
        acc = m_mul(acc, x[k]);
    }
    uint32_t inv = m_inv(acc);
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        if (i0 + k < n) out[i0 + k] = m_mul(inv, pre[k]);
        inv = m_mul(inv, x[k]);
    }
}
__global__ void __launch_bounds__(256) batch_inverse_qm31_kernel(Coords4 out, CCoords4 in, size_t n, int* flag) {
    pdl_wait();
    const size_t i0 = 4 * (blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    if (i0 >= n) return;
    QM31 x[4], pre[4];
    QM31 acc = q_from_m(1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[k] = (i0 + k < n) ? q_make(in.p[0][i0 + k], in.p[1][i0 + k], in.p[2][i0 + k], in.p[3][i0 + k]) : q_from_m(1);
        if (q_eq(x[k], q_zero())) {
            atomicExch(flag, 1);
            x[k] = q_from_m(1);
        }
        pre[k] = acc;
        acc = q_mul(acc, x[k]);
    }
    QM31 inv = q_inv(acc);
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        if (i0 + k < n) {
            const QM31 r = q_mul(inv, pre[k]);
            out.p[0][i0 + k] = r.a.a;
            out.p[1][i0 + k] = r.a.b;
            out.p[2][i0 + k] = r.b.a;
            out.p[3][i0 + k] = r.b.b;
        }
        inv = q_mul(inv, x[k]);
    }
}
cudaError_t batch_inverse_m31(uint32_t* out, const uint32_t* in, size_t n, int* d_flag, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    launch_k(batch_inverse_m31_kernel, (unsigned)((n + 1023) / 1024), 256, 0, stream, out, in, n, d_flag);
    return cudaGetLastError();
}
cudaError_t batch_inverse_qm31(uint32_t* const out[4], const uint32_t* const in[4], size_t n, int* d_flag, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    Coords4 o;
    CCoords4 s;
    for (int c = 0; c < 4; ++c) {
        o.p[c] = out[c];
        s.p[c] = in[c];
    }
    launch_k(batch_inverse_qm31_kernel, (unsigned)((n + 1023) / 1024), 256, 0, stream, o, s, n, d_flag);
    return cudaGetLastError();
}

__global__ void gather_words_kernel(uint32_t* out, const uint32_t* const* addrs, int n) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = addrs[i] ? *addrs[i] : 0u;  // a null address = a word another rank owns (sharded prover)
}
cudaError_t gather_words(uint32_t* d_out, const uint32_t* const* d_addrs, int n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    launch_k(gather_words_kernel, (n + 255) / 256, 256, 0, stream, d_out, d_addrs, n);
    return cudaGetLastError();
}

// Row-major table -> padded columns.  A CTA takes 256 consecutive rows: the row-major block is one contiguous run of
// 256 * n_cols words (read fully coalesced into shared memory, row stride padded to an odd word count so the column reads
// below are bank-conflict free), then every column is written as 256 contiguous words.
struct PadRow {
    uint32_t v[MAX_MAIN_COLS];
};
constexpr int TP_ROWS = 256;
__global__ void __launch_bounds__(TP_ROWS) transpose_pad_kernel(uint32_t* __restrict__ cols, size_t stride,
                                                                const uint32_t* __restrict__ rows, uint64_t n_rows, int n_cols,
                                                                uint64_t n_padded, const PadRow pad) {
    pdl_wait();
    extern __shared__ uint32_t tp_sm[];
    const int ncp = n_cols | 1;
    const uint64_t r0 = (uint64_t)blockIdx.x * TP_ROWS;
    const uint64_t have = r0 < n_rows ? min((uint64_t)TP_ROWS, n_rows - r0) : 0;  // real rows in this block
    const uint32_t words = (uint32_t)have * n_cols;
    const uint32_t* src = rows + r0 * n_cols;
    // the block's rows are one contiguous run of `words` words: 128-bit loads, (row, column) of each word kept incrementally
    // (one division per thread instead of one per word)
    const uint32_t nvec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (words >> 2) : 0;
    if (nvec) {
        const uint32_t dr = (4 * TP_ROWS) / n_cols, dc = (4 * TP_ROWS) - dr * n_cols;
        uint32_t r = (4 * threadIdx.x) / n_cols, c = 4 * threadIdx.x - r * n_cols;
        for (uint32_t v = threadIdx.x; v < nvec; v += TP_ROWS) {
            const uint4 x = reinterpret_cast<const uint4*>(src)[v];
            uint32_t rr = r, cc = c;
            tp_sm[rr * ncp + cc] = x.x;
            if (++cc == (uint32_t)n_cols) { cc = 0; ++rr; }
            tp_sm[rr * ncp + cc] = x.y;
            if (++cc == (uint32_t)n_cols) { cc = 0; ++rr; }
            tp_sm[rr * ncp + cc] = x.z;
            if (++cc == (uint32_t)n_cols) { cc = 0; ++rr; }
            tp_sm[rr * ncp + cc] = x.w;
            r += dr;
            c += dc;
            if (c >= (uint32_t)n_cols) { c -= n_cols; ++r; }
        }
    }
    for (uint32_t w = 4 * nvec + threadIdx.x; w < words; w += TP_ROWS) {
        uint32_t r = w / n_cols, c = w - r * n_cols;
        tp_sm[r * ncp + c] = src[w];
    }
    __syncthreads();
    const uint64_t r = r0 + threadIdx.x;
    if (r >= n_padded) return;
    const bool real = threadIdx.x < have;
    for (int c = 0; c < n_cols; ++c) cols[(size_t)c * stride + r] = real ? tp_sm[threadIdx.x * ncp + c] : pad.v[c];
}

cudaError_t transpose_pad(uint32_t* d_cols, size_t stride, const uint32_t* d_rows, uint64_t n_rows, int n_cols, int log,
                          int kind, cudaStream_t stream) {
    if (n_cols > MAX_MAIN_COLS) return cudaErrorInvalidValue;
    PadRow pad{};
    for (int c = 0; c < n_cols; ++c) pad.v[c] = padding_value(kind, c);
    uint64_t n = (uint64_t)1 << log;
    size_t smem = (size_t)TP_ROWS * (n_cols | 1) * sizeof(uint32_t);
    launch_k(transpose_pad_kernel, (unsigned)((n + TP_ROWS - 1) / TP_ROWS), TP_ROWS, smem, stream, d_cols, stride, d_rows, n_rows, n_cols, n, pad);
    return cudaGetLastError();
}

}  // namespace lb
