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
#include "launch.cuh"
#include "cfft.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
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
    pdl_wait();
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
    // X[k] lives at entry offset 2^(k-1), Y[k] at y_off + 2^k: every table is aligned to its own
    // size, so the per-thread twiddle vectors of the fast kernels are 16-byte aligned.
    size_t nx = ((size_t)1 << K);
    size_t ny = ((size_t)2 << K);
    size_t total = nx + ny;
    e = cudaMalloc(&tw->fwd, total * sizeof(uint2));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&tw->inv, total * sizeof(uint2));
    if (e != cudaSuccess) return e;
    tw->y_off = nx;
    for (int k = 0; k <= K; ++k) {
        size_t yo = nx + ((size_t)1 << k);
        uint32_t cy = 1u << k;
        launch_k(gen_twiddles_kernel, (cy + 255) / 256, 256, 0, stream, tw->fwd + yo, tw->inv + yo, k, 1);
        if (k == 0) continue;
        size_t xo = ((size_t)1 << (k - 1));
        uint32_t cx = 1u << (k - 1);
        launch_k(gen_twiddles_kernel, (cx + 255) / 256, 256, 0, stream, tw->fwd + xo, tw->inv + xo, k, 0);
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
        return base + tw->y_off + ((size_t)1 << k);
    }
    int k = n - i;  // X[n-i]
    return base + ((size_t)1 << (k - 1));
}

// ------------------------------------------------------------------------------------
// Butterflies.  P = 2^31-1 leaves no headroom in 32 bits (2P+1 = 2^32-1), so values are kept
// as arbitrary u32 representatives and folded with red() right where a sum could overflow.
// Both butterflies are closed on the full u32 range.
// ------------------------------------------------------------------------------------
// Pipe balance (B200: IMAD-class on the FMA pipe, LOP3/SHF/LEA/IADD3 on the ALU pipe, each one
// warp-instruction per 2 cycles per SM sub-partition): red() as (x & P) + (x >> 31) is two ALU
// instructions; red_fma() computes the same value as x - (x >> 31) * P with IMAD.HI + IMAD, the
// multipliers coming from kernel parameters so ptxas cannot strength-reduce them back to shifts.
#ifndef LB_BFLY_VARIANT
#define LB_BFLY_VARIANT 8
#endif
struct RedK {
    uint32_t two, neg_p;  // 2, -P (mod 2^32)
};
// red_mix: the same value as red() computed as x - (x >> 31) * P: one shift on the ALU pipe, one IMAD on the FMA pipe
__device__ __forceinline__ uint32_t red_mix(uint32_t x, RedK k) {
    uint32_t r;
    asm("{\n\t.reg .u32 h;\n\tshr.u32 h, %1, 31;\n\tmad.lo.u32 %0, h, %2, %1;\n\t}" : "=r"(r) : "r"(x), "r"(k.neg_p));
    return r;
}
template <int SLOT, int PR>
__device__ __forceinline__ uint32_t red_sel(uint32_t x, RedK k) {
    // SLOT 0: the "a" operand, SLOT 1: the multiplied / second operand; PR: butterfly index within the layer
    if (LB_BFLY_VARIANT == 4) return (SLOT == 1 && (PR & 1)) ? red_mix(x, k) : red(x);
    if (LB_BFLY_VARIANT == 5) return SLOT == 1 ? red_mix(x, k) : red(x);
    if (LB_BFLY_VARIANT == 6) return (SLOT == 1 && (PR & 3) == 0) ? red_mix(x, k) : red(x);
    if (LB_BFLY_VARIANT == 7) return red_mix(x, k);
    if (LB_BFLY_VARIANT == 8) return min(x, x - P);  // one VIADDMNMX.U32: any u32 -> [0, P+1]
    return red(x);
}
template <int PR = 0>
__device__ __forceinline__ void bfly_fwd(uint32_t& v0, uint32_t& v1, uint2 w, RedK k = RedK{2u, 0u - P}) {
    uint32_t a = red_sel<0, PR>(v0, k);                 // [0, P+1]
    uint32_t t = red_sel<1, PR>(mul_shoup(v1, w), k);   // [0, P]
    v0 = a + t;                           // <= 2P+1
    v1 = a + P - t;                       // <= 2P+1
}
template <int PR = 0>
__device__ __forceinline__ void bfly_inv(uint32_t& v0, uint32_t& v1, uint2 w, RedK k = RedK{2u, 0u - P}) {
    uint32_t a = red_sel<0, PR>(v0, k), b = red_sel<1, PR>(v1, k);  // [0, P] for inputs <= 2P
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
    size_t src_mask;       // ANDed into the source index of a high pass's first round: all ones, or 2^log_src - 1 when the
                           // top layer of a blow-up-2 evaluate is taken as the duplication it is (see cfft_evaluate_scatter)
    uint2 scale;
    RedK redk;             // {2, -P}: opaque multipliers for red_fma
    // row-sharded scatter of the final (low) pass of an evaluate: tile rows [s*R/W, (s+1)*R/W) of column c go to
    // peer[s] + (peer_col0 + c) * (R/W) + local row  (peer[] are device pointers, local or NVLink peer-mapped)
    uint32_t* peer[8];
    int n_peers;           // 0: plain store to dst
    int peer_tile_shift;   // log2(tiles per rank) = log_n - log2(W) - 12
    size_t peer_col0;      // global index of this launch's column 0
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
    pdl_wait();
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
// Fast path (log_n >= 12): fully unrolled rounds, one 16-element group per thread, every
// shared-memory offset a compile-time constant, first/last round of a pass straight from/to
// global memory (no staging), vectorised twiddle loads.
//
//   low  pass <M>: 4096 contiguous elements, layers 0..M-1 (8 <= M <= 12), 256 threads
//   high pass <M>: 2^M elements at stride 2^i_lo x W adjacent offsets (4 <= M <= 10),
//                  W = 2^(12-M) for M <= 8 (256 threads), W = 16 for M = 9, 10
// Round r of a pass works on the 4-bit index field at bit A = min(4r, M-4) and executes the
// layers of that field not done yet: field bits [4r - A, 4).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int low_pad(int e) { return e + ((e >> 5) << 2) + ((e >> 8) << 4); }
constexpr int LOW_SMEM_WORDS = 4096 + 512 + 256;

template <int W>
__device__ __forceinline__ int high_word(int e, int w) {
    return e * W + w + (W == 16 ? ((e >> 4) << 4) : 0);
}
template <int M, int W>
__host__ __device__ constexpr int high_smem_words() {
    return (1 << M) * W + (W == 16 ? (1 << M) : 0);
}

template <int CNT>
__device__ __forceinline__ void load_tw(uint2 (&w)[8], const uint2* p) {
    if (CNT == 1) {
        w[0] = __ldg(p);
    } else {
#pragma unroll
        for (int q = 0; q < CNT / 2; ++q) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + q);
            w[2 * q] = make_uint2(v.x, v.y);
            w[2 * q + 1] = make_uint2(v.z, v.w);
        }
    }
}

template <bool FWD, int PR>
__device__ __forceinline__ void bfly_at(uint32_t (&v)[16], const uint2 (&w)[8], int b, RedK k) {
    const int j0 = ((PR >> b) << (b + 1)) | (PR & ((1 << b) - 1));
    const int j1 = j0 | (1 << b);
    if (FWD)
        bfly_fwd<PR>(v[j0], v[j1], w[PR >> b], k);
    else
        bfly_inv<PR>(v[j0], v[j1], w[PR >> b], k);
}
// the 8 butterflies of one layer on 16 register-resident elements (b is a compile-time constant after unrolling)
template <bool FWD>
__device__ __forceinline__ void layer8(uint32_t (&v)[16], const uint2 (&w)[8], int b, RedK k) {
    bfly_at<FWD, 0>(v, w, b, k);
    bfly_at<FWD, 1>(v, w, b, k);
    bfly_at<FWD, 2>(v, w, b, k);
    bfly_at<FWD, 3>(v, w, b, k);
    bfly_at<FWD, 4>(v, w, b, k);
    bfly_at<FWD, 5>(v, w, b, k);
    bfly_at<FWD, 6>(v, w, b, k);
    bfly_at<FWD, 7>(v, w, b, k);
}

// layers of one index field: field bits [BLO, 4); A = bit position of the field in the tile
// index; twiddle pointer for field bit b: tw[A+b] + (tile_h << (TS-1-A-b)) + (e0_hi << (3-b))
template <bool FWD, int A, int BLO, int BHI, int TS, int NC = 1>
__device__ __forceinline__ void field_layers_range_nc(uint32_t (&v)[NC][16], const PassParams& p, uint32_t tile_h, int e0_hi) {
#pragma unroll
    for (int bb = 0; bb < BHI - BLO; ++bb) {
        const int b = FWD ? (BHI - 1 - bb) : (BLO + bb);
        const int bl = A + b;
        const uint2* twp = p.tw[bl] + ((size_t)tile_h << (TS - 1 - bl)) + ((size_t)e0_hi << (3 - b));
        uint2 w[8];
        if (b == 0) load_tw<8>(w, twp);
        if (b == 1) load_tw<4>(w, twp);
        if (b == 2) load_tw<2>(w, twp);
        if (b == 3) load_tw<1>(w, twp);
        // the twiddles depend on the position only: one load serves all NC columns held by this thread
#pragma unroll
        for (int c = 0; c < NC; ++c) layer8<FWD>(v[c], w, b, p.redk);
    }
}
template <bool FWD, int A, int BLO, int BHI, int TS>
__device__ __forceinline__ void field_layers_range(uint32_t (&v)[16], const PassParams& p, uint32_t tile_h, int e0_hi) {
    field_layers_range_nc<FWD, A, BLO, BHI, TS, 1>(reinterpret_cast<uint32_t (&)[1][16]>(v), p, tile_h, e0_hi);
}

__device__ __forceinline__ uint32_t finalize(uint32_t x, int mode, uint2 scale) {
    if (mode == 1) return canon(x);
    if (mode == 2) return canon2(mul_shoup(x, scale));
    return x;
}
// the same with the mode known at compile time: the callers branch once per round (warp-uniform) instead of once per element
template <int MODE>
__device__ __forceinline__ uint32_t finalize_m(uint32_t x, uint2 scale) {
    if (MODE == 1) return canon(x);
    if (MODE == 2) return canon2(mul_shoup(x, scale));
    return x;
}
#ifndef LB_PEER_STAGE
#define LB_PEER_STAGE 1  // fused all-to-all: send tiles to the owner rank through a shared-memory stage (coalesced NVLink stores)
#endif
#ifndef LB_UNSWITCH_FINAL
#define LB_UNSWITCH_FINAL 1
#endif

template <int M, int R>
struct RoundGeom {
    static constexpr int A = (4 * R < M - 4) ? 4 * R : (M - 4);
    static constexpr int BLO = 4 * R - A;
};

// ---- low pass ---------------------------------------------------------------------------
template <bool FWD, int M, int R, int NC>
__device__ __forceinline__ void low_round(uint32_t* sm, const PassParams& p, uint32_t tile, const uint32_t* src,
                                          uint32_t* dst, bool first, bool last, int col = 0) {
    constexpr int A = RoundGeom<M, R>::A;
    constexpr int BLO = RoundGeom<M, R>::BLO;
    const int g = threadIdx.x;
    const int e0_hi = g >> A;
    const int e0 = (e0_hi << (A + 4)) | (g & ((1 << A) - 1));
    const size_t gbase = ((size_t)tile << 12) + e0;
    uint32_t v[NC][16];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const uint32_t* srcc = src + (size_t)c * p.src_stride;
        uint32_t* smc = sm + c * LOW_SMEM_WORDS;
        if (first) {
            if (A == 0) {
                const size_t n_src = (size_t)1 << p.log_src;  // zero extension beyond n_src
                if (gbase + 16 <= n_src) {
                    const uint4* s4 = reinterpret_cast<const uint4*>(srcc + gbase);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 t = s4[q];
                        v[c][4 * q] = t.x; v[c][4 * q + 1] = t.y; v[c][4 * q + 2] = t.z; v[c][4 * q + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c][j] = gbase + j < n_src ? srcc[gbase + j] : 0u;
                }
            } else if (p.log_src >= p.log_n) {
                // full-size source (always the case when this is not the only pass): no zero-extension checks
#pragma unroll
                for (int j = 0; j < 16; ++j) v[c][j] = srcc[gbase + ((size_t)j << A)];
            } else {
                const size_t n_src = (size_t)1 << p.log_src;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    size_t gi = gbase + ((size_t)j << A);
                    v[c][j] = gi < n_src ? srcc[gi] : 0u;
                }
            }
        } else {
            const int sb = low_pad(e0);
            if (A == 0) {
                const uint4* s4 = reinterpret_cast<const uint4*>(smc + sb);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 t = s4[q];
                    v[c][4 * q] = t.x; v[c][4 * q + 1] = t.y; v[c][4 * q + 2] = t.z; v[c][4 * q + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[c][j] = smc[sb + low_pad(j << A)];
            }
        }
    }
    field_layers_range_nc<FWD, A, BLO, 4, 12, NC>(v, p, tile, e0_hi);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        uint32_t* dstc = dst + (size_t)c * p.dst_stride;
        uint32_t* smc = sm + c * LOW_SMEM_WORDS;
        if (last) {
            if (A == 0 && FWD && LB_PEER_STAGE && p.n_peers > 0) {
                // fused all-to-all, staged: park the finished tile in shared memory; the copy loop below sends it to the owner
                // rank with warp-contiguous 512-byte stores (a thread's own 64-byte run would reach NVLink as 16-byte writes)
                uint4* s4 = reinterpret_cast<uint4*>(smc + low_pad(e0));
#pragma unroll
                for (int q = 0; q < 4; ++q) s4[q] = make_uint4(v[c][4 * q], v[c][4 * q + 1], v[c][4 * q + 2], v[c][4 * q + 3]);
            } else if (A == 0) {
                uint4* d4 = reinterpret_cast<uint4*>(dstc + gbase);
                if (FWD && p.n_peers > 0) {
                    // fused all-to-all: this tile belongs to the row shard of rank `tile >> peer_tile_shift`; store it straight
                    // into that rank's receive buffer (NVLink peer memory), already in [column][local row] layout
                    const uint32_t owner = tile >> p.peer_tile_shift;
                    const size_t rows_local = (size_t)4096 << p.peer_tile_shift;
                    const size_t local_row = ((size_t)(tile & ((1u << p.peer_tile_shift) - 1)) << 12) + e0;
                    d4 = reinterpret_cast<uint4*>(p.peer[owner] + (p.peer_col0 + col + c) * rows_local + local_row);
                }
#if LB_UNSWITCH_FINAL
                // forward: the low pass is the last one (canonical store); inverse: last only for single-pass sizes (scale)
                auto st4 = [&](auto mode_tag) {
                    constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        d4[q] = make_uint4(finalize_m<MODE>(v[c][4 * q], p.scale), finalize_m<MODE>(v[c][4 * q + 1], p.scale),
                                           finalize_m<MODE>(v[c][4 * q + 2], p.scale), finalize_m<MODE>(v[c][4 * q + 3], p.scale));
                };
                if (p.final_mode == 0) st4(std::integral_constant<int, 0>{});
                else if (p.final_mode == 1) st4(std::integral_constant<int, 1>{});
                else st4(std::integral_constant<int, 2>{});
#else
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    d4[q] = make_uint4(finalize(v[c][4 * q], p.final_mode, p.scale), finalize(v[c][4 * q + 1], p.final_mode, p.scale),
                                       finalize(v[c][4 * q + 2], p.final_mode, p.scale), finalize(v[c][4 * q + 3], p.final_mode, p.scale));
#endif
            } else {
                // inverse: the low pass is the first pass (lazy store) unless it is the only one (scale by 1/n)
                auto st = [&](auto mode_tag) {
                    constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
                    for (int j = 0; j < 16; ++j) dstc[gbase + ((size_t)j << A)] = finalize_m<MODE>(v[c][j], p.scale);
                };
                if (p.final_mode == 0) st(std::integral_constant<int, 0>{});
                else if (p.final_mode == 1) st(std::integral_constant<int, 1>{});
                else st(std::integral_constant<int, 2>{});
            }
        } else {
            const int sb = low_pad(e0);
            if (A == 0) {
                uint4* d4 = reinterpret_cast<uint4*>(smc + sb);
#pragma unroll
                for (int q = 0; q < 4; ++q) d4[q] = make_uint4(v[c][4 * q], v[c][4 * q + 1], v[c][4 * q + 2], v[c][4 * q + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) smc[sb + low_pad(j << A)] = v[c][j];
            }
        }
    }
    if (last && A == 0 && FWD && LB_PEER_STAGE && p.n_peers > 0) {
        __syncthreads();
        const uint32_t owner = tile >> p.peer_tile_shift;
        const size_t rows_local = (size_t)4096 << p.peer_tile_shift;
        const size_t row0 = (size_t)(tile & ((1u << p.peer_tile_shift) - 1)) << 12;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const uint32_t* smc = sm + c * LOW_SMEM_WORDS;
            uint4* d4 = reinterpret_cast<uint4*>(p.peer[owner] + (p.peer_col0 + col + c) * rows_local + row0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int chunk = threadIdx.x + 256 * k;  // 16-byte chunk of the 4096-element tile
                uint4 t = *reinterpret_cast<const uint4*>(smc + low_pad(4 * chunk));
                d4[chunk] = make_uint4(finalize(t.x, p.final_mode, p.scale), finalize(t.y, p.final_mode, p.scale),
                                       finalize(t.z, p.final_mode, p.scale), finalize(t.w, p.final_mode, p.scale));
            }
        }
    }
}

#ifndef LB_LOW_MINB
#define LB_LOW_MINB 6
#endif
#ifndef LB_LOW_MINB2
#define LB_LOW_MINB2 4
#endif
#ifndef LB_LOW_NC
#define LB_LOW_NC 2
#endif
#ifndef LB_HIGH_MINB
#define LB_HIGH_MINB 4
#endif
#ifndef LB_CPB_MAX
#define LB_CPB_MAX 1
#endif
// NC columns per thread: the NC tiles (same position, adjacent columns) go through the rounds together, so every
// twiddle load and index computation is shared by NC butterflies and each warp has NC x the independent work.
template <bool FWD, int M, int NC>
__global__ void __launch_bounds__(256, NC == 1 ? LB_LOW_MINB : LB_LOW_MINB2) cfft_low_fast(PassParams p, int cols_per_block) {
    pdl_wait();
    extern __shared__ __align__(16) uint32_t sm[];  // NC * LOW_SMEM_WORDS
    constexpr int NR = (M + 3) / 4;
    const uint32_t tile = blockIdx.x;
    const int c0 = blockIdx.y * cols_per_block * NC;
    const int c1 = min(p.n_cols, c0 + cols_per_block * NC);
    for (int c = c0; c < c1; c += NC) {
        const uint32_t* src = p.src + (size_t)c * p.src_stride;
        uint32_t* dst = p.dst + (size_t)c * p.dst_stride;
        if (FWD) {
            if constexpr (NR == 3) { low_round<FWD, M, 2, NC>(sm, p, tile, src, dst, true, false, c); __syncthreads(); }
            low_round<FWD, M, 1, NC>(sm, p, tile, src, dst, NR == 2, false, c);
            __syncthreads();
            low_round<FWD, M, 0, NC>(sm, p, tile, src, dst, false, true, c);
        } else {
            low_round<FWD, M, 0, NC>(sm, p, tile, src, dst, true, false, c);
            __syncthreads();
            low_round<FWD, M, 1, NC>(sm, p, tile, src, dst, false, NR == 2, c);
            if constexpr (NR == 3) { __syncthreads(); low_round<FWD, M, 2, NC>(sm, p, tile, src, dst, false, true, c); }
        }
        __syncthreads();
    }
}

// ---- high pass --------------------------------------------------------------------------
// ILO: compile-time i_lo (0 = runtime) so the 16 strided global offsets become immediates.
// ZEXT (forward first pass only): 0 = source is full size; 1 = source is the lower half
// (blow-up 2: the upper half is zero, the top layer degenerates to a copy); 2 = generic bound check.
template <bool FWD, int M, int W, int R, int ILO, int ZEXT, int NC>
__device__ __forceinline__ void high_round(uint32_t* sm, const PassParams& p, uint32_t tile_h, size_t g_base,
                                           const uint32_t* src, uint32_t* dst, bool first, bool last) {
    constexpr int A = RoundGeom<M, R>::A;
    constexpr int BLO = RoundGeom<M, R>::BLO;
    constexpr bool TOP = (A + 4 == M);
    constexpr int SMW = high_smem_words<M, W>();
    const int ilo = ILO ? ILO : p.i_lo;
    const int w = threadIdx.x % W;
    const int g = threadIdx.x / W;
    const int e0_hi = g >> A;
    const int e0 = (e0_hi << (A + 4)) | (g & ((1 << A) - 1));
    const size_t gb = g_base + ((size_t)e0 << ilo) + w;
    const int sb = high_word<W>(e0, w);
    uint32_t v[NC][16];
    bool half_zero = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (first) {
            const uint32_t* sp = src + (size_t)c * p.src_stride + (gb & p.src_mask);
            if (ZEXT == 2) {
                const size_t n_src = (size_t)1 << p.log_src;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    size_t off = (size_t)(j << A) << ilo;
                    v[c][j] = gb + off < n_src ? sp[off] : 0u;
                }
            } else if (ZEXT == 1 && TOP) {
                // tile covers the top layers (tile_h == 0): elements with the top index bit set are zero
#pragma unroll
                for (int j = 0; j < 8; ++j) v[c][j] = sp[(size_t)(j << A) << ilo];
#pragma unroll
                for (int j = 8; j < 16; ++j) v[c][j] = v[c][j - 8];
                half_zero = true;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[c][j] = sp[(size_t)(j << A) << ilo];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[c][j] = sm[c * SMW + sb + high_word<W>(j << A, 0)];
        }
    }
    if (ZEXT == 1 && TOP && FWD && BLO < 4) {
        // top layer on (v, 0) pairs is the identity copy done above; run the remaining layers
        if (half_zero) {
            field_layers_range_nc<FWD, A, BLO, 3, M, NC>(v, p, tile_h, e0_hi);
        } else {
            field_layers_range_nc<FWD, A, BLO, 4, M, NC>(v, p, tile_h, e0_hi);
        }
    } else {
        field_layers_range_nc<FWD, A, BLO, 4, M, NC>(v, p, tile_h, e0_hi);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (last) {
            uint32_t* dp = dst + (size_t)c * p.dst_stride + gb;
#if LB_UNSWITCH_FINAL
            auto st = [&](auto mode_tag) {
                constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
                for (int j = 0; j < 16; ++j) dp[(size_t)(j << A) << ilo] = finalize_m<MODE>(v[c][j], p.scale);
            };
            // forward high passes are never the last pass (lazy store); inverse: the top pass scales by 1/n
            if (FWD || p.final_mode == 0) st(std::integral_constant<int, 0>{});
            else st(std::integral_constant<int, 2>{});
#else
#pragma unroll
            for (int j = 0; j < 16; ++j) dp[(size_t)(j << A) << ilo] = finalize(v[c][j], p.final_mode, p.scale);
#endif
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) sm[c * SMW + sb + high_word<W>(j << A, 0)] = v[c][j];
        }
    }
}

template <int M>
struct HighGeom {
    static constexpr int W = (M <= 8) ? (1 << (12 - M)) : 16;
    static constexpr int THREADS = (1 << (M - 4)) * W;
    static constexpr int MIN_BLOCKS = (M <= 8) ? LB_HIGH_MINB : 1;
};

#ifndef LB_HIGH_NC
#define LB_HIGH_NC 1  // measured: two columns per thread pay off in the low pass (twiddle reuse), not here
#endif
#ifndef LB_HIGH_MINB2
#define LB_HIGH_MINB2 2
#endif
template <bool FWD, int M, int ILO, int ZEXT, int NC>
__global__ void __launch_bounds__(HighGeom<M>::THREADS, NC == 1 ? HighGeom<M>::MIN_BLOCKS : ((M <= 8) ? LB_HIGH_MINB2 : 1))
    cfft_high_fast(PassParams p, int cols_per_block) {
    pdl_wait();
    constexpr int W = HighGeom<M>::W;
    constexpr int NR = (M + 3) / 4;
    extern __shared__ __align__(16) uint32_t smh[];
    const int ilo = ILO ? ILO : p.i_lo;
    const uint32_t l_tiles = (1u << ilo) / W;
    const uint32_t tile_h = blockIdx.x / l_tiles, lt = blockIdx.x % l_tiles;
    const size_t g_base = ((size_t)tile_h << (ilo + M)) + (size_t)lt * W;
    const int c0 = blockIdx.y * cols_per_block * NC;
    const int c1 = min(p.n_cols, c0 + cols_per_block * NC);
    for (int c = c0; c < c1; c += NC) {
        const uint32_t* src = p.src + (size_t)c * p.src_stride;
        uint32_t* dst = p.dst + (size_t)c * p.dst_stride;
        if constexpr (NR == 1) {
            high_round<FWD, M, W, 0, ILO, ZEXT, NC>(smh, p, tile_h, g_base, src, dst, true, true);
        } else if constexpr (FWD) {
            if constexpr (NR == 3) { high_round<FWD, M, W, 2, ILO, ZEXT, NC>(smh, p, tile_h, g_base, src, dst, true, false); __syncthreads(); }
            high_round<FWD, M, W, 1, ILO, ZEXT, NC>(smh, p, tile_h, g_base, src, dst, NR == 2, false);
            __syncthreads();
            high_round<FWD, M, W, 0, ILO, ZEXT, NC>(smh, p, tile_h, g_base, src, dst, false, true);
            __syncthreads();
        } else {
            high_round<FWD, M, W, 0, ILO, 0, NC>(smh, p, tile_h, g_base, src, dst, true, false);
            __syncthreads();
            high_round<FWD, M, W, 1, ILO, 0, NC>(smh, p, tile_h, g_base, src, dst, false, NR == 2);
            if constexpr (NR == 3) { __syncthreads(); high_round<FWD, M, W, 2, ILO, 0, NC>(smh, p, tile_h, g_base, src, dst, false, true); }
            __syncthreads();
        }
    }
}

// ---- high pass, vectorised along the offset dimension ---------------------------------------------------
// Each thread owns VW adjacent offsets (one 64- or 128-bit access) of its 16 strided rows, i.e. VW independent radix-16
// groups that share every twiddle (twiddles depend on the row only).  Versus the scalar high pass this divides the
// load/store, address and twiddle instructions per butterfly by VW; a tile is 2^M rows x 16 VW offsets.  VW = 4 executes 21 %
// fewer instructions than VW = 1 but needs 128 registers; VW = 2 keeps twice the resident warps and is the faster one.
// (The same idea applied to the low pass - 64 elements per thread, all accesses 128-bit - was measured slower, 0.757 ms
// per round trip, for the same reason, and is not kept.)
#ifndef LB_HIGH_VEC
#define LB_HIGH_VEC 1
#endif
#ifndef LB_HIGH_VEC_MAX_M
#define LB_HIGH_VEC_MAX_M 10  // 8: only the n = 20 high pass is vectorised; 9 / 10: also the single high pass of n = 21 / 22
#endif
#ifndef LB_HIGH_VW
#define LB_HIGH_VW 2  // measured: 2 -> 0.592 ms round trip, 4 -> 0.605 ms (128 registers, half the resident warps)
#endif
template <int VW>
struct VecIO;
template <>
struct VecIO<4> {
    __device__ static __forceinline__ void load(const uint32_t* p, uint32_t (&o)[4]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
    __device__ static __forceinline__ void store(uint32_t* p, const uint32_t (&o)[4]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
    }
};
template <>
struct VecIO<2> {
    __device__ static __forceinline__ void load(const uint32_t* p, uint32_t (&o)[2]) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        o[0] = t.x; o[1] = t.y;
    }
    __device__ static __forceinline__ void store(uint32_t* p, const uint32_t (&o)[2]) {
        *reinterpret_cast<uint2*>(p) = make_uint2(o[0], o[1]);
    }
};

template <bool FWD, int M, int R, int ILO, int ZEXT, int VW>
__device__ __forceinline__ void high_round_vec(uint32_t* sm, const PassParams& p, uint32_t tile_h, size_t g_base,
                                               const uint32_t* src, uint32_t* dst, bool first, bool last) {
    constexpr int A = RoundGeom<M, R>::A;
    constexpr int BLO = RoundGeom<M, R>::BLO;
    constexpr bool TOP = (A + 4 == M);
    constexpr int WT = 16, W = WT * VW;
    const int wt = threadIdx.x % WT;
    const int g = threadIdx.x / WT;
    const int e0_hi = g >> A;
    const int e0 = (e0_hi << (A + 4)) | (g & ((1 << A) - 1));
    const size_t gb = g_base + ((size_t)e0 << ILO) + VW * wt;
    const int sb = e0 * W + VW * wt;
    uint32_t v[VW][16];
    constexpr bool HALF = (ZEXT == 1 && TOP && FWD);  // the upper half of the rows is the zero extension
    if (first) {
        const uint32_t* sp = src + (gb & p.src_mask);
#pragma unroll
        for (int j = 0; j < (HALF ? 8 : 16); ++j) {
            uint32_t t[VW];
            VecIO<VW>::load(sp + ((size_t)(j << A) << ILO), t);
#pragma unroll
            for (int k = 0; k < VW; ++k) v[k][j] = t[k];
        }
        if (HALF) {
#pragma unroll
            for (int j = 8; j < 16; ++j)
#pragma unroll
                for (int k = 0; k < VW; ++k) v[k][j] = v[k][j - 8];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t t[VW];
            VecIO<VW>::load(sm + sb + (j << A) * W, t);
#pragma unroll
            for (int k = 0; k < VW; ++k) v[k][j] = t[k];
        }
    }
    if (HALF && first && BLO < 4)
        field_layers_range_nc<FWD, A, BLO, 3, M, VW>(v, p, tile_h, e0_hi);  // the top layer on (v, 0) pairs is the copy above
    else
        field_layers_range_nc<FWD, A, BLO, 4, M, VW>(v, p, tile_h, e0_hi);
    if (last) {
        uint32_t* dp = dst + gb;
        auto st = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t t[VW];
#pragma unroll
                for (int k = 0; k < VW; ++k) t[k] = finalize_m<MODE>(v[k][j], p.scale);
                VecIO<VW>::store(dp + ((size_t)(j << A) << ILO), t);
            }
        };
        if (FWD || p.final_mode == 0) st(std::integral_constant<int, 0>{});
        else st(std::integral_constant<int, 2>{});
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t t[VW];
#pragma unroll
            for (int k = 0; k < VW; ++k) t[k] = v[k][j];
            VecIO<VW>::store(sm + sb + (j << A) * W, t);
        }
    }
}

template <bool FWD, int M, int ILO, int ZEXT, int VW>
__global__ void __launch_bounds__((1 << (M - 4)) * 16, M == 8 ? (VW == 4 ? 2 : 4) : (M == 9 && VW == 2 ? 2 : 1)) cfft_high_vec(PassParams p) {
    pdl_wait();
    static_assert(M >= 8 && M <= 10, "two or three rounds of up to four layers");
    extern __shared__ __align__(16) uint32_t smv[];
    constexpr uint32_t W = 16 * VW;
    constexpr uint32_t l_tiles = (1u << ILO) / W;
    const uint32_t tile_h = blockIdx.x / l_tiles, lt = blockIdx.x % l_tiles;
    const size_t g_base = ((size_t)tile_h << (ILO + M)) + (size_t)lt * W;
    const uint32_t* src = p.src + (size_t)blockIdx.y * p.src_stride;
    uint32_t* dst = p.dst + (size_t)blockIdx.y * p.dst_stride;
    constexpr bool THREE = (M > 8);  // M = 9, 10: a third round for the one or two layers above the first eight
    if constexpr (FWD) {
        if constexpr (THREE) {
            high_round_vec<FWD, M, 2, ILO, ZEXT, VW>(smv, p, tile_h, g_base, src, dst, true, false);
            __syncthreads();
        }
        high_round_vec<FWD, M, 1, ILO, ZEXT, VW>(smv, p, tile_h, g_base, src, dst, !THREE, false);
        __syncthreads();
        high_round_vec<FWD, M, 0, ILO, ZEXT, VW>(smv, p, tile_h, g_base, src, dst, false, true);
    } else {
        high_round_vec<FWD, M, 0, ILO, 0, VW>(smv, p, tile_h, g_base, src, dst, true, false);
        __syncthreads();
        high_round_vec<FWD, M, 1, ILO, 0, VW>(smv, p, tile_h, g_base, src, dst, false, !THREE);
        if constexpr (THREE) {
            __syncthreads();
            high_round_vec<FWD, M, 2, ILO, 0, VW>(smv, p, tile_h, g_base, src, dst, false, true);
        }
    }
}

template <bool FWD, int M, int ILO, int ZEXT>
static cudaError_t launch_high_vec(const PassParams& p, cudaStream_t stream) {
    constexpr int VW = LB_HIGH_VW;
    constexpr size_t W = 16 * VW;
    size_t tiles = ((size_t)1 << (p.log_n - ILO - M)) * (((size_t)1 << ILO) / W);
    constexpr size_t smem = ((size_t)1 << M) * W * sizeof(uint32_t);
    if (p.n_cols > 65535) return cudaErrorInvalidValue;
    auto k = cfft_high_vec<FWD, M, ILO, ZEXT, VW>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)tiles, (unsigned)p.n_cols);
    launch_k(k, grid, (1 << (M - 4)) * 16, smem, stream, p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// Host-side pass planning
// ------------------------------------------------------------------------------------
static constexpr int LOW_TS_MAX = 12;

struct Plan {
    int n_pass;
    int i_lo[8], m[8];
    bool fast;
};

static Plan make_plan(int n) {
    Plan pl{};
    pl.i_lo[0] = 0;
    if (n < LOW_TS_MAX) {  // small transforms: one generic tile covers the column
        pl.n_pass = 1;
        pl.m[0] = n;
        pl.fast = false;
        return pl;
    }
    pl.fast = true;
    if (n == 12) {
        pl.n_pass = 1;
        pl.m[0] = 12;
        return pl;
    }
    if (n <= 15) {  // low n-4 (9..11) + high 4
        pl.n_pass = 2;
        pl.m[0] = n - 4;
        pl.i_lo[1] = n - 4;
        pl.m[1] = 4;
        return pl;
    }
    int rem = n - 12;
    int n_high = (rem + 9) / 10;
    pl.n_pass = 1 + n_high;
    pl.m[0] = 12;
    int at = 12;
    for (int k = 0; k < n_high; ++k) {
        int mk = rem / n_high + (k < rem % n_high ? 1 : 0);
        pl.i_lo[1 + k] = at;
        pl.m[1 + k] = mk;
        at += mk;
    }
    return pl;
}

// passes for the lowest `layers` (>= 16) layers of a transform: the low pass and as many high passes as needed
static Plan make_plan_layers(int layers) { return make_plan(layers); }

#ifndef LB_LDE_DUP
#define LB_LDE_DUP 1
#endif

static int pick_cols_per_block(size_t tiles, int n_cols, int sm_count) {
    int cpb = 1;
    size_t target_blocks = (size_t)sm_count * 16;
    while (cpb < n_cols && cpb < LB_CPB_MAX && tiles * ((n_cols + cpb - 1) / cpb) > target_blocks) cpb *= 2;
    return cpb;
}

template <bool FWD>
static cudaError_t launch_generic(const PassParams& p, int sm_count, cudaStream_t stream) {
    size_t tiles = (size_t)1 << (p.log_n - p.ts);
    size_t el = (size_t)1 << p.ts;
    size_t smem = (el + (el >> 4) + 1) * sizeof(uint32_t);
    int cpb = pick_cols_per_block(tiles, p.n_cols, sm_count);
    dim3 grid((unsigned)tiles, (unsigned)((p.n_cols + cpb - 1) / cpb));
    if (grid.y > 65535) return cudaErrorInvalidValue;
    launch_k(cfft_pass_kernel<FWD, 1>, grid, 256, smem, stream, p, cpb);
    return cudaGetLastError();
}

template <bool FWD, int M, int NC>
static cudaError_t launch_low_nc(const PassParams& p, int sm_count, cudaStream_t stream) {
    size_t tiles = (size_t)1 << (p.log_n - 12);
    int groups = p.n_cols / NC;  // the caller passes a multiple of NC
    int cpb = pick_cols_per_block(tiles, groups, sm_count);
    dim3 grid((unsigned)tiles, (unsigned)((groups + cpb - 1) / cpb));
    if (grid.y > 65535) return cudaErrorInvalidValue;
    constexpr size_t smem = (size_t)NC * LOW_SMEM_WORDS * sizeof(uint32_t);
    auto k = cfft_low_fast<FWD, M, NC>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    launch_k(k, grid, 256, smem, stream, p, cpb);
    return cudaGetLastError();
}

template <bool FWD, int M>
static cudaError_t launch_low_m(const PassParams& p, int sm_count, cudaStream_t stream) {
    constexpr int NC = LB_LOW_NC;
    int main_cols = (p.n_cols / NC) * NC;
    if (NC > 1 && main_cols > 0) {
        PassParams q = p;
        q.n_cols = main_cols;
        cudaError_t e = launch_low_nc<FWD, M, NC>(q, sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    int rest = (NC > 1) ? p.n_cols - main_cols : p.n_cols;
    if (rest > 0) {
        PassParams q = p;
        int skip = (NC > 1) ? main_cols : 0;
        q.src = p.src + (size_t)skip * p.src_stride;
        q.dst = p.dst + (size_t)skip * p.dst_stride;
        q.peer_col0 = p.peer_col0 + skip;
        q.n_cols = rest;
        return launch_low_nc<FWD, M, 1>(q, sm_count, stream);
    }
    return cudaSuccess;
}

template <bool FWD>
static cudaError_t launch_low(const PassParams& p, int sm_count, cudaStream_t stream) {
    switch (p.m) {
        case 8: return launch_low_m<FWD, 8>(p, sm_count, stream);
        case 9: return launch_low_m<FWD, 9>(p, sm_count, stream);
        case 10: return launch_low_m<FWD, 10>(p, sm_count, stream);
        case 11: return launch_low_m<FWD, 11>(p, sm_count, stream);
        case 12: return launch_low_m<FWD, 12>(p, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

template <bool FWD, int M, int ILO, int ZEXT, int NC>
static cudaError_t launch_high_nc(const PassParams& p, int sm_count, cudaStream_t stream) {
    constexpr int W = HighGeom<M>::W;
    if (((size_t)1 << p.i_lo) < (size_t)W) return cudaErrorInvalidValue;
    size_t tiles = ((size_t)1 << (p.log_n - p.i_lo - M)) * (((size_t)1 << p.i_lo) / W);
    size_t smem = (size_t)NC * high_smem_words<M, W>() * sizeof(uint32_t);
    int groups = p.n_cols / NC;
    int cpb = pick_cols_per_block(tiles, groups, sm_count);
    dim3 grid((unsigned)tiles, (unsigned)((groups + cpb - 1) / cpb));
    if (grid.y > 65535) return cudaErrorInvalidValue;
    auto k = cfft_high_fast<FWD, M, ILO, ZEXT, NC>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    launch_k(k, grid, HighGeom<M>::THREADS, smem, stream, p, cpb);
    return cudaGetLastError();
}

template <bool FWD, int M, int ILO, int ZEXT>
static cudaError_t launch_high_k(const PassParams& p, int sm_count, cudaStream_t stream) {
    // two columns per thread where the doubled tile still leaves room for several CTAs per SM
    constexpr int NC = (M <= 8) ? LB_HIGH_NC : 1;
    int main_cols = (p.n_cols / NC) * NC;
    if (NC > 1 && main_cols > 0) {
        PassParams q = p;
        q.n_cols = main_cols;
        cudaError_t e = launch_high_nc<FWD, M, ILO, ZEXT, NC>(q, sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    int rest = (NC > 1) ? p.n_cols - main_cols : p.n_cols;
    if (rest > 0) {
        PassParams q = p;
        int skip = (NC > 1) ? main_cols : 0;
        q.src = p.src + (size_t)skip * p.src_stride;
        q.dst = p.dst + (size_t)skip * p.dst_stride;
        q.n_cols = rest;
        return launch_high_nc<FWD, M, ILO, ZEXT, 1>(q, sm_count, stream);
    }
    return cudaSuccess;
}

template <bool FWD, int M>
static cudaError_t launch_high_m(const PassParams& p, int sm_count, cudaStream_t stream) {
    // zero extension only matters for the forward pass that reads the coefficients
    int zext = 0;
    if (FWD && p.log_src < p.log_n) {
        bool top = (p.i_lo + M == p.log_n);
        zext = (top && p.log_src == p.log_n - 1 && M >= 4) ? 1 : 2;
    }
    if constexpr (M >= 8 && M <= LB_HIGH_VEC_MAX_M && LB_HIGH_VEC) {
        // the vectorised kernel needs whole 64-offset tiles below the pass (i_lo = 12) and no ragged zero extension
        if (p.i_lo == 12 && zext != 2) {
            if (FWD && zext == 1) return launch_high_vec<FWD, M, 12, 1>(p, stream);
            return launch_high_vec<FWD, M, 12, 0>(p, stream);
        }
    }
    if constexpr (FWD) {
        if (p.i_lo == 12) {
            if (zext == 1) return launch_high_k<FWD, M, 12, 1>(p, sm_count, stream);
            if (zext == 2) return launch_high_k<FWD, M, 0, 2>(p, sm_count, stream);
            return launch_high_k<FWD, M, 12, 0>(p, sm_count, stream);
        }
        if (zext) return launch_high_k<FWD, M, 0, 2>(p, sm_count, stream);
        return launch_high_k<FWD, M, 0, 0>(p, sm_count, stream);
    } else {
        if (p.i_lo == 12) return launch_high_k<FWD, M, 12, 0>(p, sm_count, stream);
        return launch_high_k<FWD, M, 0, 0>(p, sm_count, stream);
    }
}

template <bool FWD>
static cudaError_t launch_high(const PassParams& p, int sm_count, cudaStream_t stream) {
    switch (p.m) {
        case 4: return launch_high_m<FWD, 4>(p, sm_count, stream);
        case 5: return launch_high_m<FWD, 5>(p, sm_count, stream);
        case 6: return launch_high_m<FWD, 6>(p, sm_count, stream);
        case 7: return launch_high_m<FWD, 7>(p, sm_count, stream);
        case 8: return launch_high_m<FWD, 8>(p, sm_count, stream);
        case 9: return launch_high_m<FWD, 9>(p, sm_count, stream);
        case 10: return launch_high_m<FWD, 10>(p, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

template <bool FWD>
static cudaError_t launch_pass(const PassParams& p, const Plan& pl, int k, int sm_count, cudaStream_t stream) {
    if (!pl.fast) return launch_generic<FWD>(p, sm_count, stream);
    return k == 0 ? launch_low<FWD>(p, sm_count, stream) : launch_high<FWD>(p, sm_count, stream);
}

// L2-resident column groups: a transform of log size >= 16 is several passes over its columns.  Run over all columns at
// once, every pass streams the whole batch through HBM (1.8x the algorithmic bytes at n = 20).  Walking the batch in groups
// whose working set fits the 126 MB L2 lets the later passes of a group read what the previous pass just wrote from L2, so
// HBM sees each element once in and once out per transform.  Measured on B200 (r2, 64 x 2^20 round trip): 0.576 ms
// ungrouped, 0.644 / 0.652 / 0.717 / 0.728 / 0.874 ms with 96 / 64 / 48 / 40 / 24 MiB groups - the passes are bound by the
// integer pipes, not by HBM, and the smaller launches lose more to their ragged last wave than L2 hits give back.  So the
// grouping is OFF by default (LB_CFFT_GROUP_MB = 0); the knob stays for batches whose passes are memory-bound.
static size_t cfft_group_bytes() {
    static long mb = -1;
    if (mb < 0) {
        const char* e = getenv("LB_CFFT_GROUP_MB");
        mb = e ? atol(e) : 0;
        if (mb < 0) mb = 0;
    }
    return (size_t)mb << 20;
}
static int cfft_group_cols(size_t bytes_per_col, int n_cols, int n_pass) {
    size_t budget = cfft_group_bytes();
    if (budget == 0 || n_pass < 2 || (size_t)n_cols * bytes_per_col <= budget) return n_cols;
    size_t g = budget / bytes_per_col;
    if (g < 2) return n_cols;  // a single column does not fit: nothing to gain
    g &= ~(size_t)1;           // the kernels take columns in pairs
    // balance the groups (e.g. 64 columns with room for 10 -> 7 groups of 10,10,10,10,8,8,8)
    int n_groups = (int)((n_cols + g - 1) / g);
    int per = (n_cols + n_groups - 1) / n_groups;
    per = (per + 1) & ~1;
    return per;
}

static cudaError_t cfft_interpolate_all(const Twiddles* tw, uint32_t* data, size_t stride, int n_cols, int log_n, int sm_count,
                                        cudaStream_t stream);

cudaError_t cfft_interpolate(const Twiddles* tw, uint32_t* data, size_t stride, int n_cols, int log_n, int sm_count,
                             cudaStream_t stream) {
    if (log_n < 1 || log_n > tw->max_log) return cudaErrorInvalidValue;
    if (n_cols == 0) return cudaSuccess;
    Plan pl = make_plan(log_n);
    int g = cfft_group_cols(sizeof(uint32_t) << log_n, n_cols, pl.n_pass);
    for (int c0 = 0; c0 < n_cols; c0 += g) {
        cudaError_t e = cfft_interpolate_all(tw, data + (size_t)c0 * stride, stride, std::min(g, n_cols - c0), log_n, sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static cudaError_t cfft_interpolate_all(const Twiddles* tw, uint32_t* data, size_t stride, int n_cols, int log_n, int sm_count,
                                        cudaStream_t stream) {
    if (log_n < 1 || log_n > tw->max_log) return cudaErrorInvalidValue;
    if (n_cols == 0) return cudaSuccess;
    Plan pl = make_plan(log_n);
    uint32_t inv_n = m_inv(1u << log_n);
    for (int k = 0; k < pl.n_pass; ++k) {
        PassParams p{};
        p.redk = RedK{2u, 0u - P};
        p.src = data;
        p.dst = data;
        p.src_stride = p.dst_stride = stride;
        p.n_cols = n_cols;
        p.log_n = log_n;
        p.log_src = log_n;
        p.i_lo = pl.i_lo[k];
        p.m = pl.m[k];
        p.ts = (k == 0) ? std::min(log_n, LOW_TS_MAX) : p.m;
        p.final_mode = (k == pl.n_pass - 1) ? 2 : 0;
        p.src_mask = ~(size_t)0;
        p.scale = make_uint2(inv_n, shoup_companion(inv_n));
        for (int b = 0; b < p.m; ++b) p.tw[b] = layer_tw(tw, true, log_n, p.i_lo + b);
        cudaError_t e = launch_pass<false>(p, pl, k, sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t cfft_evaluate(const Twiddles* tw, const uint32_t* coeffs, size_t src_stride, int log_in, uint32_t* out,
                          size_t dst_stride, int log_out, int n_cols, int sm_count, cudaStream_t stream) {
    return cfft_evaluate_scatter(tw, coeffs, src_stride, log_in, out, dst_stride, log_out, n_cols, nullptr, 0, 0, sm_count, stream);
}

static cudaError_t cfft_evaluate_scatter_all(const Twiddles* tw, const uint32_t* coeffs, size_t src_stride, int log_in, uint32_t* out,
                                             size_t dst_stride, int log_out, int n_cols, uint32_t* const* peers, int n_peers,
                                             size_t peer_col0, int sm_count, cudaStream_t stream);

cudaError_t cfft_evaluate_scatter(const Twiddles* tw, const uint32_t* coeffs, size_t src_stride, int log_in, uint32_t* out,
                                  size_t dst_stride, int log_out, int n_cols, uint32_t* const* peers, int n_peers,
                                  size_t peer_col0, int sm_count, cudaStream_t stream) {
    if (log_out < 1 || log_out > tw->max_log || log_in > log_out || log_in < 0) return cudaErrorInvalidValue;
    if (n_cols == 0) return cudaSuccess;
    const bool dup = LB_LDE_DUP && log_in == log_out - 1 && log_out >= 17 && coeffs != out;
    Plan pl = dup ? make_plan_layers(log_out - 1) : make_plan(log_out);
    // working set of a column: its source (when it is not the destination) plus its destination
    size_t per_col = (sizeof(uint32_t) << log_out) + (coeffs != out ? (sizeof(uint32_t) << log_in) : 0);
    int g = cfft_group_cols(per_col, n_cols, pl.n_pass);
    for (int c0 = 0; c0 < n_cols; c0 += g) {
        cudaError_t e = cfft_evaluate_scatter_all(tw, coeffs + (size_t)c0 * src_stride, src_stride, log_in, out + (size_t)c0 * dst_stride,
                                                  dst_stride, log_out, std::min(g, n_cols - c0), peers, n_peers, peer_col0 + c0,
                                                  sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static cudaError_t cfft_evaluate_scatter_all(const Twiddles* tw, const uint32_t* coeffs, size_t src_stride, int log_in, uint32_t* out,
                                             size_t dst_stride, int log_out, int n_cols, uint32_t* const* peers, int n_peers,
                                             size_t peer_col0, int sm_count, cudaStream_t stream) {
    if (log_out < 1 || log_out > tw->max_log || log_in > log_out || log_in < 0) return cudaErrorInvalidValue;
    if (n_cols == 0) return cudaSuccess;
    int log_w = 0;
    if (n_peers > 0) {
        if (n_peers > 8 || (n_peers & (n_peers - 1)) || !peers) return cudaErrorInvalidValue;
        while ((1 << log_w) < n_peers) ++log_w;
        if (log_out - log_w < 12 || log_out < 16) return cudaErrorInvalidValue;  // whole 4096-row tiles per rank, multi-pass plan
    }
    // Blow-up 2 (log_in = log_out - 1, the commit path): the top layer pairs every coefficient with a zero, i.e. it only
    // copies the lower half into the upper half.  The plan then covers layers 0 .. log_out-2 and the top-most pass reads
    // its source at (index mod 2^log_in): one layer and, for log_out = 23, one whole HBM pass less.
    const bool dup = LB_LDE_DUP && log_in == log_out - 1 && log_out >= 17 && coeffs != out;
    Plan pl = dup ? make_plan_layers(log_out - 1) : make_plan(log_out);
    for (int k = pl.n_pass - 1; k >= 0; --k) {
        PassParams p{};
        p.redk = RedK{2u, 0u - P};
        bool first = (k == pl.n_pass - 1);
        p.src = first ? coeffs : out;
        p.dst = out;
        p.src_stride = first ? src_stride : dst_stride;
        p.dst_stride = dst_stride;
        p.n_cols = n_cols;
        p.log_n = log_out;
        p.log_src = (first && !dup) ? log_in : log_out;
        p.src_mask = (first && dup) ? (((size_t)1 << log_in) - 1) : ~(size_t)0;
        p.i_lo = pl.i_lo[k];
        p.m = pl.m[k];
        p.ts = (k == 0) ? std::min(log_out, LOW_TS_MAX) : p.m;
        p.final_mode = (k == 0) ? 1 : 0;
        if (k == 0 && n_peers > 0) {
            p.n_peers = n_peers;
            for (int r = 0; r < n_peers; ++r) p.peer[r] = peers[r];
            p.peer_tile_shift = log_out - log_w - 12;
            p.peer_col0 = peer_col0;
        }
        for (int b = 0; b < p.m; ++b) p.tw[b] = layer_tw(tw, false, log_out, p.i_lo + b);
        cudaError_t e = launch_pass<true>(p, pl, k, sm_count, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// stwo-format twiddle tree of a root half coset half_odds(root_log): per layer, x of the first
// half of the k-times doubled coset, bit-reversed; concatenated; padded with 1.
__global__ void export_stwo_twiddles_kernel(const uint2* fwd_x_base, uint32_t* out, int root_log) {
    pdl_wait();
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
    const uint2* x = fwd_x_base + ((size_t)1 << (kx - 1));
    out[t] = x[off_in_layer].x;
}

cudaError_t twiddles_export_stwo(const Twiddles* tw, int root_log, uint32_t* d_out, cudaStream_t stream) {
    if (root_log < 1 || root_log > tw->max_log - 1) return cudaErrorInvalidValue;
    size_t total = (size_t)1 << root_log;
    launch_k(export_stwo_twiddles_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, tw->fwd, d_out, root_log);
    return cudaGetLastError();
}

}  // namespace lb
