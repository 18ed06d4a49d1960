// Circle-FFT over M31 for sm_100a: twiddle generation, interpolate (values -> coefficients),
// evaluate / low-degree extension (coefficients -> values on a larger canonic coset).
//
// Replaces stwo's SimdBackend PolyOps::{precompute_twiddles, interpolate, evaluate} as reached
// from /root/reference/crates/prover/src/prover.rs:38-42 (twiddles), every
// tree_builder.extend_evals (e.g. crates/air/src/components/add/witness.rs:51,164 -> interpolate)
// and every tree_builder.commit (prover.rs:59,179,298 -> evaluate on the blown-up domain).
//
// Data layout: one column = 2^n contiguous u32 in HBM, bit-reversed circle-domain order;
// a batch = n_cols columns at a fixed stride.  Layer i pairs elements at distance 2^i with
// twiddle index h = idx >> (i+1)  (layer 0: y-twiddles, layers >= 1: x-twiddles).
//
// Kernel structure (one kernel template, two tilings):
//   W == 1  "low" pass : a tile is 2^ts contiguous elements of one column (layers 0..m-1)
//   W == 16 "high" pass: a tile is 2^m elements at stride 2^i_lo x 16 adjacent offsets, so a
//                        warp always touches 64-byte contiguous segments
// Inside a tile the m layers run as rounds of <= 4 layers on 16 register-resident elements per
// thread; rounds exchange through shared memory.  Butterflies are lazy: values live in
// [0, 2^32) and are folded only where the next operation needs it (see m31.cuh).
#include "cfft.cuh"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace lb {

// ------------------------------------------------------------------------------------
// Twiddles
// ------------------------------------------------------------------------------------
__constant__ Pt c_gen_pow2[31];  // G * 2^j

__device__ __forceinline__ Pt point_of_index(uint32_t idx) {
    Pt r = {1, 0};
    bool first = true;
#pragma unroll 1
    for (int j = 0; j < 31; ++j) {
        if ((idx >> j) & 1) {
            if (first) {
                r = c_gen_pow2[j];
                first = false;
            } else {
                r = pt_add(r, c_gen_pow2[j]);
            }
        }
    }
    return r;
}

// X[k][h] = x(half_odds(k).at(bitrev(h, k-1))), h < 2^(k-1)      (line layers)
// Y[k][h] = y(half_odds(k).at(bitrev(h, k))),   h < 2^k          (circle layer)
__global__ void gen_twiddles_kernel(uint2* fwd, uint2* inv, int k, int is_y) {
    uint32_t bits = is_y ? k : k - 1;
    uint32_t n = 1u << bits;
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    uint32_t i = bits ? (__brev(h) >> (32 - bits)) : 0;
    uint32_t init = 1u << (31 - (k + 2));
    uint32_t step = 1u << (31 - k);
    uint32_t idx = (init + i * step) & 0x7FFFFFFFu;
    Pt p = point_of_index(idx);
    uint32_t t = is_y ? p.y : p.x;
    uint32_t ti = m_inv(t);
    fwd[h] = make_uint2(t, shoup_companion(t));
    inv[h] = make_uint2(ti, shoup_companion(ti));
}

static void host_gen_pow2(Pt* out) {
    Pt p = {2, 1268011823u};
    for (int j = 0; j < 31; ++j) {
        out[j] = p;
        p = pt_add(p, p);
    }
}

cudaError_t twiddles_create(Twiddles* tw, int max_log, cudaStream_t stream) {
    // supports domains of log size <= max_log:  X[1..max_log-1], Y[0..max_log-1]
    tw->max_log = max_log;
    Pt hp[31];
    host_gen_pow2(hp);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_gen_pow2, hp, sizeof(hp), 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    int K = max_log - 1;
    if (K < 1) K = 1;
    size_t nx = ((size_t)1 << K);      // sum_{k=1..K} 2^(k-1) = 2^K - 1
    size_t ny = ((size_t)2 << K);      // sum_{k=0..K} 2^k = 2^(K+1) - 1
    size_t total = nx + ny;
    e = cudaMalloc(&tw->fwd, total * sizeof(uint2));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&tw->inv, total * sizeof(uint2));
    if (e != cudaSuccess) return e;
    tw->y_off = nx;
    for (int k = 0; k <= K; ++k) {
        size_t yo = nx + ((size_t)1 << k) - 1;
        uint32_t cy = 1u << k;
        gen_twiddles_kernel<<<(cy + 255) / 256, 256, 0, stream>>>(tw->fwd + yo, tw->inv + yo, k, 1);
        if (k == 0) continue;
        size_t xo = ((size_t)1 << (k - 1)) - 1;
        uint32_t cx = 1u << (k - 1);
        gen_twiddles_kernel<<<(cx + 255) / 256, 256, 0, stream>>>(tw->fwd + xo, tw->inv + xo, k, 0);
    }
    return cudaGetLastError();
}

void twiddles_destroy(Twiddles* tw) {
    if (tw->fwd) cudaFree(tw->fwd);
    if (tw->inv) cudaFree(tw->inv);
    tw->fwd = tw->inv = nullptr;
}

// twiddle array of layer i for a domain of log size n
static inline const uint2* layer_tw(const Twiddles* tw, bool inverse, int n, int i) {
    const uint2* base = inverse ? tw->inv : tw->fwd;
    if (i == 0) {
        int k = n - 1;  // Y[n-1]
        return base + tw->y_off + ((size_t)1 << k) - 1;
    }
    int k = n - i;  // X[n-i]
    return base + ((size_t)1 << (k - 1)) - 1;
}

// ------------------------------------------------------------------------------------
// Butterflies.  P = 2^31-1 leaves no headroom in 32 bits (2P+1 = 2^32-1), so values are kept
// as arbitrary u32 representatives and folded with red() right where a sum could overflow.
// Both butterflies are closed on the full u32 range.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void bfly_fwd(uint32_t& v0, uint32_t& v1, uint2 w) {
    uint32_t a = red(v0);                 // [0, P+1]
    uint32_t t = red(mul_shoup(v1, w));   // [0, P]
    v0 = a + t;                           // <= 2P+1
    v1 = a + P - t;                       // <= 2P+1
}
__device__ __forceinline__ void bfly_inv(uint32_t& v0, uint32_t& v1, uint2 w) {
    uint32_t a = red(v0), b = red(v1);  // [0, P] for inputs <= 2P
    v0 = a + b;                         // [0, 2P]
    v1 = mul_shoup(a + P - b, w);       // [0, 2P)
}

struct PassParams {
    const uint32_t* src;   // read in the first round
    uint32_t* dst;         // written in the last round (may alias src)
    size_t src_stride, dst_stride;  // column strides (elements)
    int n_cols;
    int log_n;             // transform size
    int log_src;           // log of the valid prefix of src (< log_n: zero extension)
    int i_lo;              // first global layer of this pass
    int m;                 // layers in this pass
    int ts;                // log2(FFT elements per tile), ts >= m
    int final_mode;        // 0 lazy store, 1 canonical, 2 scale (interpolate) + canonical
    uint2 scale;
    const uint2* tw[12];   // twiddle arrays of layers i_lo .. i_lo+m-1
};

template <int W>
__device__ __forceinline__ int sm_index(int e, int w) {
    if (W == 1) return e + (e >> 4);
    return e * W + w;
}

// One round: L layers [a, a+L) of the tile on 2^L register elements per work item.
template <bool FWD, int L, int W>
__device__ __forceinline__ void do_round(uint32_t* sm, const PassParams& p, int a, uint32_t tile_h, int tid, int nthreads) {
    constexpr int R = 1 << L;
    const int n_groups = 1 << (p.ts - L);
    const int n_items = n_groups * W;
    for (int item = tid; item < n_items; item += nthreads) {
        int w = (W == 1) ? 0 : (item % W);
        int g = (W == 1) ? item : (item / W);
        int g_hi = g >> a, g_lo = g & ((1 << a) - 1);
        int e0 = (g_hi << (a + L)) | g_lo;
        uint32_t v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = sm[sm_index<W>(e0 | (j << a), w)];
#pragma unroll
        for (int bb = 0; bb < L; ++bb) {
            const int b = FWD ? (L - 1 - bb) : bb;
            const int bl = a + b;  // local layer
            // twiddle index base: (tile_h << (ts-1-bl)) + (g_hi << (L-1-b)) + (j >> (b+1))
            const uint2* twp = p.tw[bl] + ((size_t)tile_h << (p.ts - 1 - bl)) + ((size_t)g_hi << (L - 1 - b));
            constexpr int dummy = 0;
            (void)dummy;
            uint2 wv[R / 2];
            const int ntw = 1 << (L - 1 - b);
#pragma unroll
            for (int q = 0; q < R / 2; ++q)
                if (q < ntw) wv[q] = __ldg(twp + q);
#pragma unroll
            for (int pr = 0; pr < R / 2; ++pr) {
                const int j0 = ((pr >> b) << (b + 1)) | (pr & ((1 << b) - 1));
                const int j1 = j0 + (1 << b);
                if (FWD)
                    bfly_fwd(v[j0], v[j1], wv[pr >> b]);
                else
                    bfly_inv(v[j0], v[j1], wv[pr >> b]);
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) sm[sm_index<W>(e0 | (j << a), w)] = v[j];
    }
}

template <bool FWD, int W>
__device__ __forceinline__ void run_round(uint32_t* sm, const PassParams& p, int a, int L, uint32_t tile_h) {
    const int tid = threadIdx.x, nt = blockDim.x;
    switch (L) {
        case 4: do_round<FWD, 4, W>(sm, p, a, tile_h, tid, nt); break;
        case 3: do_round<FWD, 3, W>(sm, p, a, tile_h, tid, nt); break;
        case 2: do_round<FWD, 2, W>(sm, p, a, tile_h, tid, nt); break;
        default: do_round<FWD, 1, W>(sm, p, a, tile_h, tid, nt); break;
    }
}

// grid: x = tile, y = column group.  Each block loops over its columns.
template <bool FWD, int W>
__global__ void __launch_bounds__(256) cfft_pass_kernel(PassParams p, int cols_per_block) {
    extern __shared__ uint32_t sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tile_elems = 1 << p.ts;
    // tile geometry
    uint32_t tile = blockIdx.x;
    uint32_t tile_h;       // upper index feeding the twiddle offset
    size_t g_base;         // global index of (e=0, w=0)
    int e_shift;           // global stride of e = 2^e_shift
    if (W == 1) {
        tile_h = tile;
        g_base = (size_t)tile << p.ts;
        e_shift = 0;
    } else {
        uint32_t l_tiles = (1u << p.i_lo) / W;
        uint32_t hi_part = tile / l_tiles, lt = tile % l_tiles;
        tile_h = hi_part;
        g_base = ((size_t)hi_part << (p.i_lo + p.m)) + (size_t)lt * W;
        e_shift = p.i_lo;
    }
    const size_t n_src = (size_t)1 << p.log_src;
    const int c0 = blockIdx.y * cols_per_block;
    const int c1 = min(p.n_cols, c0 + cols_per_block);
    for (int c = c0; c < c1; ++c) {
        const uint32_t* src = p.src + (size_t)c * p.src_stride;
        uint32_t* dst = p.dst + (size_t)c * p.dst_stride;
        // ---- stage in
        for (int t = tid; t < tile_elems * W; t += nt) {
            int w = (W == 1) ? 0 : (t % W);
            int e = (W == 1) ? t : (t / W);
            size_t gi = g_base + ((size_t)e << e_shift) + w;
            sm[sm_index<W>(e, w)] = gi < n_src ? src[gi] : 0u;
        }
        __syncthreads();
        // ---- rounds: layer groups aligned at multiples of 4 (short round on top)
        const int n_rounds = (p.m + 3) >> 2;
        for (int r = 0; r < n_rounds; ++r) {
            const int rr = FWD ? (n_rounds - 1 - r) : r;
            const int a = rr << 2;
            const int L = min(4, p.m - a);
            run_round<FWD, W>(sm, p, a, L, tile_h);
            __syncthreads();
        }
        // ---- stage out
        for (int t = tid; t < tile_elems * W; t += nt) {
            int w = (W == 1) ? 0 : (t % W);
            int e = (W == 1) ? t : (t / W);
            size_t gi = g_base + ((size_t)e << e_shift) + w;
            uint32_t x = sm[sm_index<W>(e, w)];
            if (p.final_mode == 1)
                x = canon(x);
            else if (p.final_mode == 2)
                x = canon2(mul_shoup(x, p.scale));
            dst[gi] = x;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// Host-side pass planning
// ------------------------------------------------------------------------------------
static constexpr int LOW_TS_MAX = 12;
static constexpr int HIGH_M_MAX = 10;
static constexpr int HIGH_W = 16;

struct Plan {
    int n_pass;
    int i_lo[8], m[8];
};

static Plan make_plan(int n) {
    Plan pl{};
    if (n <= LOW_TS_MAX) {
        pl.n_pass = 1;
        pl.i_lo[0] = 0;
        pl.m[0] = n;
        return pl;
    }
    // low pass takes at least half of the layers (high tiles need i_lo >= 4), at most 12
    int low = std::min(LOW_TS_MAX, std::max(n - HIGH_M_MAX, (n + 1) / 2));
    int rem = n - low;
    int n_high = (rem + HIGH_M_MAX - 1) / HIGH_M_MAX;
    pl.n_pass = 1 + n_high;
    pl.i_lo[0] = 0;
    pl.m[0] = low;
    int at = low;
    for (int k = 0; k < n_high; ++k) {
        int mk = rem / n_high + (k < rem % n_high ? 1 : 0);
        pl.i_lo[1 + k] = at;
        pl.m[1 + k] = mk;
        at += mk;
    }
    return pl;
}

template <bool FWD>
static cudaError_t launch_pass(const PassParams& p, bool low, int sm_count, cudaStream_t stream) {
    size_t tiles, smem;
    if (low) {
        tiles = (size_t)1 << (p.log_n - p.ts);
        size_t el = (size_t)1 << p.ts;
        smem = (el + (el >> 4) + 1) * sizeof(uint32_t);
    } else {
        tiles = ((size_t)1 << (p.log_n - p.i_lo - p.m)) * (((size_t)1 << p.i_lo) / HIGH_W);
        smem = ((size_t)1 << p.ts) * HIGH_W * sizeof(uint32_t);
    }
    // choose columns per block so that the grid has a few waves of blocks
    int cols_per_block = 1;
    size_t target_blocks = (size_t)sm_count * 8;
    while (cols_per_block < p.n_cols && tiles * ((p.n_cols + cols_per_block - 1) / cols_per_block) > 2 * target_blocks &&
           cols_per_block < 8)
        cols_per_block *= 2;
    dim3 grid((unsigned)tiles, (unsigned)((p.n_cols + cols_per_block - 1) / cols_per_block));
    if (grid.y > 65535) return cudaErrorInvalidValue;
    cudaError_t e;
    if (low) {
        auto k = cfft_pass_kernel<FWD, 1>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k<<<grid, 256, smem, stream>>>(p, cols_per_block);
    } else {
        auto k = cfft_pass_kernel<FWD, HIGH_W>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k<<<grid, 256, smem, stream>>>(p, cols_per_block);
    }
    return cudaGetLastError();
}

cudaError_t cfft_interpolate(const Twiddles* tw, uint32_t* data, size_t stride, int n_cols, int log_n, int sm_count,
                             cudaStream_t stream) {
    if (log_n < 1 || log_n > tw->max_log) return cudaErrorInvalidValue;
    if (n_cols == 0) return cudaSuccess;
    Plan pl = make_plan(log_n);
    uint32_t inv_n = m_inv(1u << log_n);
    for (int k = 0; k < pl.n_pass; ++k) {
        PassParams p{};
        p.src = data;
        p.dst = data;
        p.src_stride = p.dst_stride = stride;
        p.n_cols = n_cols;
        p.log_n = log_n;
        p.log_src = log_n;
        p.i_lo = pl.i_lo[k];
        p.m = pl.m[k];
        bool low = (k == 0);
        p.ts = low ? std::min(log_n, std::max(p.m, LOW_TS_MAX)) : p.m;
        if (low && p.ts > log_n) p.ts = log_n;
        p.final_mode = (k == pl.n_pass - 1) ? 2 : 0;
        p.scale = make_uint2(inv_n, shoup_companion(inv_n));
        for (int b = 0; b < p.m; ++b) p.tw[b] = layer_tw(tw, true, log_n, p.i_lo + b);
        cudaError_t e = launch_pass<false>(p, low, sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t cfft_evaluate(const Twiddles* tw, const uint32_t* coeffs, size_t src_stride, int log_in, uint32_t* out,
                          size_t dst_stride, int log_out, int n_cols, int sm_count, cudaStream_t stream) {
    if (log_out < 1 || log_out > tw->max_log || log_in > log_out || log_in < 0) return cudaErrorInvalidValue;
    if (n_cols == 0) return cudaSuccess;
    Plan pl = make_plan(log_out);
    for (int k = pl.n_pass - 1; k >= 0; --k) {
        PassParams p{};
        bool first = (k == pl.n_pass - 1);
        p.src = first ? coeffs : out;
        p.dst = out;
        p.src_stride = first ? src_stride : dst_stride;
        p.dst_stride = dst_stride;
        p.n_cols = n_cols;
        p.log_n = log_out;
        p.log_src = first ? log_in : log_out;
        p.i_lo = pl.i_lo[k];
        p.m = pl.m[k];
        bool low = (k == 0);
        p.ts = low ? std::min(log_out, std::max(p.m, LOW_TS_MAX)) : p.m;
        p.final_mode = (k == 0) ? 1 : 0;
        for (int b = 0; b < p.m; ++b) p.tw[b] = layer_tw(tw, false, log_out, p.i_lo + b);
        cudaError_t e = launch_pass<true>(p, low, sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// stwo-format twiddle tree of a root half coset half_odds(root_log): per layer, x of the first
// half of the k-times doubled coset, bit-reversed; concatenated; padded with 1.
__global__ void export_stwo_twiddles_kernel(const uint2* fwd_x_base, uint32_t* out, int root_log) {
    // layer k (0-based) = X[root_log - k], size 2^(root_log-k-1), at out offset sum of previous sizes
    size_t total = ((size_t)1 << root_log);
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    if (t == total - 1) {
        out[t] = 1;
        return;
    }
    // find layer: offsets o_k = 2^root_log - 2^(root_log-k)
    size_t rem = total - 1 - t;  // in [1, total-1]
    int lg = 63 - __clzll((long long)rem);   // rem in [2^lg, 2^(lg+1))
    int kx = lg + 1;                         // X[kx] has 2^(kx-1) = 2^lg entries
    size_t off_in_layer = ((size_t)2 << lg) - 1 - rem;  // position inside the layer
    const uint2* x = fwd_x_base + (((size_t)1 << (kx - 1)) - 1);
    out[t] = x[off_in_layer].x;
}

cudaError_t twiddles_export_stwo(const Twiddles* tw, int root_log, uint32_t* d_out, cudaStream_t stream) {
    if (root_log < 1 || root_log > tw->max_log - 1) return cudaErrorInvalidValue;
    size_t total = (size_t)1 << root_log;
    export_stwo_twiddles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(tw->fwd, d_out, root_log);
    return cudaGetLastError();
}

}  // namespace lb
