// Blake2s-256 Merkle layers for sm_100a.
//
// Replaces stwo MerkleOps<Blake2sMerkleHasher>::commit_on_layer as reached from every
// tree_builder.commit(channel) (/root/reference/crates/prover/src/prover.rs:59,179,298) and
// from the FRI layer commitments inside stwo::prover::prove (prover.rs:311-312).
//
//   node(i) = Blake2s-256( left(32B) || right(32B) || col_0[i] || col_1[i] || ... )   LE u32 words
// (children absent on the deepest layer).  One thread hashes one node; columns are read
// column-major so a warp reads 128 contiguous bytes per column; digests are stored as 8 u32
// per node so the next layer reads its two children as one 64-byte segment.
#include "launch.cuh"
#include "merkle.cuh"

#include "blake2s.cuh"

namespace lb {

#ifndef LB_MERKLE_SUBTREE_LATENCY_FORM
#define LB_MERKLE_SUBTREE_LATENCY_FORM 1  // upper levels of merkle_subtree_kernel: plain-add compression (few warps per CTA)
#endif
#ifndef LB_MERKLE_FMA
#define LB_MERKLE_FMA 1
#endif

__device__ __forceinline__ void compress_dev(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t f0, uint32_t one) {
#if LB_MERKLE_FMA
    blake2s_compress_fma(h, m, t0, 0, f0, one);
#else
    (void)one;
    blake2s_compress(h, m, t0, 0, f0);
#endif
}

// hash of node i of a layer: children (prev != nullptr) then the layer's column values.
// SAME_KERNEL: the children were written earlier by this kernel (another thread of the CTA, before a block barrier):
// they are read with ld.global.cg so the loads can never be served by the non-coherent read-only path.
template <bool SAME_KERNEL = false>
__device__ __forceinline__ void hash_node(uint32_t h[8], const uint32_t* prev, const uint32_t* const* __restrict__ cols,
                                          int n_cols, uint32_t i, uint32_t one) {
    blake2s_init(h);
    uint32_t m[16];
    uint32_t t = 0;
    int c = 0;
    bool have_block = false;
    if (prev) {
        const uint4* pp = reinterpret_cast<const uint4*>(prev + (size_t)i * 16);
        uint4 a, b, cc, d;
        if (SAME_KERNEL) {
            a = __ldcg(pp); b = __ldcg(pp + 1); cc = __ldcg(pp + 2); d = __ldcg(pp + 3);
        } else {
            a = pp[0]; b = pp[1]; cc = pp[2]; d = pp[3];
        }
        m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
        m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
        m[8] = cc.x; m[9] = cc.y; m[10] = cc.z; m[11] = cc.w;
        m[12] = d.x; m[13] = d.y; m[14] = d.z; m[15] = d.w;
        t = 64;
        have_block = true;
    }
    while (true) {
        if (have_block) {
            bool last = (c >= n_cols);
            compress_dev(h, m, t, last ? 0xFFFFFFFFu : 0u, one);
            if (last) break;
        }
        // build next block from columns c .. c+15
        int take = min(16, n_cols - c);
        if (take <= 0) {
            // nothing at all was hashed yet: Blake2s("") = one zero block, t = 0, final
#pragma unroll
            for (int j = 0; j < 16; ++j) m[j] = 0;
            have_block = true;
            c = n_cols;  // forces last
            continue;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) m[j] = (j < take) ? cols[c + j][i] : 0u;
        c += take;
        t += 4u * take;
        have_block = true;
    }
}

// cols: device array of n_cols column base pointers (each 2^log_size u32).
__global__ void __launch_bounds__(256) merkle_layer_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ prev,
                                                           const uint32_t* const* __restrict__ cols, int n_cols,
                                                           uint32_t n_nodes, uint32_t one) {
    pdl_wait();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t h[8];
    hash_node(h, prev, cols, n_cols, i, one);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 8);
    o[0] = make_uint4(h[0], h[1], h[2], h[3]);
    o[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

cudaError_t merkle_commit_layer(uint32_t* out, const uint32_t* prev, const uint32_t* const* d_cols, int n_cols,
                                int log_size, cudaStream_t stream) {
    uint32_t n = 1u << log_size;
    // (two nodes per thread, as in merkle_layer_small_kernel<2>, was measured slower here: the pointer-table loads and 96
    // registers cost more than the second chain gains - 12.53 vs 12.23 ms on the 61-column proof)
    launch_k(merkle_layer_kernel, (n + 255) / 256, 256, 0, stream, out, prev, d_cols, n_cols, n, 1u);
    return cudaGetLastError();
}

// Same, for layers with at most MERKLE_SMALL_COLS columns: the column pointers travel in the kernel parameters, so
// no pointer table has to be staged in device memory first (the FRI layer trees: 4 coordinate columns each).
// NH nodes per thread (256 apart, so every access stays coalesced): Blake2s is one dependent chain of 4-wide G steps,
// and a second independent compression per thread lets the scheduler keep both integer pipes busy - measured with
// scripts/ubench/b2s.cu, register-only: 23.4 -> 26.8 G compressions/s.
template <int NH>
__global__ void __launch_bounds__(256, NH == 1 ? 4 : 2) merkle_layer_small_kernel(uint32_t* __restrict__ out,
                                                                                  const uint32_t* __restrict__ prev,
                                                                                  const __grid_constant__ MerkleColsArg cols,
                                                                                  int n_cols, uint32_t n_nodes, uint32_t one) {
    pdl_wait();
    const uint32_t i0 = blockIdx.x * (256 * NH) + threadIdx.x;
    if (i0 >= n_nodes) return;  // NH > 1 is only launched for n_nodes that are multiples of 256 * NH
    uint32_t h[NH][8], m[NH][16];
    uint32_t t = 0;
#pragma unroll
    for (int q = 0; q < NH; ++q) blake2s_init(h[q]);
    if (prev) {
#pragma unroll
        for (int q = 0; q < NH; ++q) {
            const uint4* pp = reinterpret_cast<const uint4*>(prev + (size_t)(i0 + 256 * q) * 16);
            uint4 a = pp[0], b = pp[1], cc = pp[2], d = pp[3];
            m[q][0] = a.x; m[q][1] = a.y; m[q][2] = a.z; m[q][3] = a.w;
            m[q][4] = b.x; m[q][5] = b.y; m[q][6] = b.z; m[q][7] = b.w;
            m[q][8] = cc.x; m[q][9] = cc.y; m[q][10] = cc.z; m[q][11] = cc.w;
            m[q][12] = d.x; m[q][13] = d.y; m[q][14] = d.z; m[q][15] = d.w;
        }
        t = 64;
#pragma unroll
        for (int q = 0; q < NH; ++q) compress_dev(h[q], m[q], t, n_cols == 0 ? 0xFFFFFFFFu : 0u, one);
    }
    if (n_cols > 0 || !prev) {
#pragma unroll
        for (int q = 0; q < NH; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) m[q][j] = (j < n_cols) ? cols.p[j][i0 + 256 * q] : 0u;
        t += 4u * n_cols;
#pragma unroll
        for (int q = 0; q < NH; ++q) compress_dev(h[q], m[q], t, 0xFFFFFFFFu, one);
    }
#pragma unroll
    for (int q = 0; q < NH; ++q) {
        uint4* o = reinterpret_cast<uint4*>(out + (size_t)(i0 + 256 * q) * 8);
        o[0] = make_uint4(h[q][0], h[q][1], h[q][2], h[q][3]);
        o[1] = make_uint4(h[q][4], h[q][5], h[q][6], h[q][7]);
    }
}

#ifndef LB_MERKLE_NH
#define LB_MERKLE_NH 2
#endif
cudaError_t merkle_commit_layer_small(uint32_t* out, const uint32_t* prev, const MerkleColsArg& cols, int n_cols,
                                      int log_size, cudaStream_t stream) {
    if (n_cols < 0 || n_cols > MERKLE_SMALL_COLS) return cudaErrorInvalidValue;
    uint32_t n = 1u << log_size;
    // two nodes per thread once the layer still fills the machine that way (148 SMs x 2 CTAs x 512 nodes)
    if (LB_MERKLE_NH == 2 && log_size >= 18)
        launch_k(merkle_layer_small_kernel<2>, n / 512, 256, 0, stream, out, prev, cols, n_cols, n, 1u);
    else
        launch_k(merkle_layer_small_kernel<1>, (n + 255) / 256, 256, 0, stream, out, prev, cols, n_cols, n, 1u);
    return cudaGetLastError();
}

// The top of a tree in one launch: given layer `from_log` (<= MERKLE_TOP_MAX_LOG) already hashed, one CTA computes the
// column-less layers from_log-1 .. 0, each into its own buffer, with a block barrier between levels.  Saves one
// ~3 us dependent launch per level on every tree (the FRI layers make ~25 trees per proof).
// These levels are a chain of dependent compressions on a handful of warps, so latency is what counts: children are handed
// from level to level through shared memory (each level is still written to its global buffer for the decommitment) and the
// compression is the plain-add form, whose three-input adds make a shorter dependent chain than the IMAD-forced throughput form.
__device__ __forceinline__ void hash_children_latency(uint32_t h[8], const uint4 x0, const uint4 x1, const uint4 x2, const uint4 x3) {
    uint32_t m[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
    blake2s_init(h);
    blake2s_compress(h, m, 64, 0, 0xFFFFFFFFu);
}

__global__ void __launch_bounds__(512) merkle_top_kernel(MerkleTopArgs a, uint32_t one) {
    pdl_wait();
    pdl_launch_dependents();
    (void)one;
    __shared__ uint4 sm[2 * 512];  // digest i of the level just hashed: sm[2 * i], sm[2 * i + 1]
    const uint32_t tid = threadIdx.x;
    for (int log = a.from_log - 1; log >= 0; --log) {
        const uint32_t n = 1u << log;
        uint4 x0, x1, x2, x3;
        if (tid < n) {
            if (log == a.from_log - 1) {  // written by the previous launch
                const uint4* pp = reinterpret_cast<const uint4*>(a.layers[log + 1] + (size_t)tid * 16);
                x0 = pp[0]; x1 = pp[1]; x2 = pp[2]; x3 = pp[3];
            } else {
                x0 = sm[4 * tid]; x1 = sm[4 * tid + 1]; x2 = sm[4 * tid + 2]; x3 = sm[4 * tid + 3];
            }
        }
        __syncthreads();  // every child has been read: the slots can be overwritten
        if (tid < n) {
            uint32_t h[8];
            hash_children_latency(h, x0, x1, x2, x3);
            const uint4 lo = make_uint4(h[0], h[1], h[2], h[3]), hi = make_uint4(h[4], h[5], h[6], h[7]);
            uint4* o = reinterpret_cast<uint4*>(a.layers[log] + (size_t)tid * 8);
            o[0] = lo;
            o[1] = hi;
            sm[2 * tid] = lo;
            sm[2 * tid + 1] = hi;
        }
        __syncthreads();
    }
}

cudaError_t merkle_commit_top(const MerkleTopArgs& args, cudaStream_t stream) {
    if (args.from_log < 1 || args.from_log > MERKLE_TOP_MAX_LOG) return cudaErrorInvalidValue;
    launch_k(merkle_top_kernel, 1, 512, 0, stream, args, 1u);
    return cudaGetLastError();
}

// ---- several levels per launch (narrow layers) ---------------------------------------------------------
// CTA b hashes nodes [b * S, (b + 1) * S) of layer log_top (S = 512: two nodes per thread on the widest level), then
// their S/2 parents, ... for `depth` levels.  Each level is written to its global buffer (the tree is kept for
// decommitment) and read back by the same CTA after a block barrier.
__global__ void __launch_bounds__(256, 2) merkle_subtree_kernel(const __grid_constant__ MerkleSubtreeArgs a, uint32_t one) {
    pdl_wait();
    const uint32_t n_top = 1u << a.log_top;
    const uint32_t S = min(512u, n_top);  // nodes of layer log_top per CTA
    const uint32_t base = blockIdx.x * S;
    {
        uint32_t h[2][8], m[2][16];
        const uint32_t i0 = base + threadIdx.x, i1 = i0 + 256;
        const bool on0 = threadIdx.x < S, on1 = threadIdx.x + 256 < S;
        const uint32_t j0 = on0 ? i0 : base, j1 = on1 ? i1 : base;  // idle lanes recompute node `base` (never stored)
        blake2s_init(h[0]);
        blake2s_init(h[1]);
        uint32_t t = 0;
        if (a.prev) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint4* pp = reinterpret_cast<const uint4*>(a.prev + (size_t)(q ? j1 : j0) * 16);
                uint4 x = pp[0], y = pp[1], z = pp[2], w = pp[3];
                m[q][0] = x.x; m[q][1] = x.y; m[q][2] = x.z; m[q][3] = x.w;
                m[q][4] = y.x; m[q][5] = y.y; m[q][6] = y.z; m[q][7] = y.w;
                m[q][8] = z.x; m[q][9] = z.y; m[q][10] = z.z; m[q][11] = z.w;
                m[q][12] = w.x; m[q][13] = w.y; m[q][14] = w.z; m[q][15] = w.w;
            }
            t = 64;
            compress_dev(h[0], m[0], t, a.n_cols == 0 ? 0xFFFFFFFFu : 0u, one);
            compress_dev(h[1], m[1], t, a.n_cols == 0 ? 0xFFFFFFFFu : 0u, one);
        }
        if (a.n_cols > 0 || !a.prev) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                m[0][j] = (j < a.n_cols) ? a.cols.p[j][j0] : 0u;
                m[1][j] = (j < a.n_cols) ? a.cols.p[j][j1] : 0u;
            }
            t += 4u * a.n_cols;
            compress_dev(h[0], m[0], t, 0xFFFFFFFFu, one);
            compress_dev(h[1], m[1], t, 0xFFFFFFFFu, one);
        }
        if (on0) {
            uint4* o = reinterpret_cast<uint4*>(a.layers[0] + (size_t)i0 * 8);
            o[0] = make_uint4(h[0][0], h[0][1], h[0][2], h[0][3]);
            o[1] = make_uint4(h[0][4], h[0][5], h[0][6], h[0][7]);
        }
        if (on1) {
            uint4* o = reinterpret_cast<uint4*>(a.layers[0] + (size_t)i1 * 8);
            o[0] = make_uint4(h[1][0], h[1][1], h[1][2], h[1][3]);
            o[1] = make_uint4(h[1][4], h[1][5], h[1][6], h[1][7]);
        }
    }
    for (int d = 1; d < a.depth; ++d) {
        __syncthreads();  // level d-1 of this CTA's range is complete (global writes are visible block-wide after the barrier)
        const uint32_t cnt = S >> d;
        if (cnt == 0) break;
        if (threadIdx.x < cnt) {
            const uint32_t i = (base >> d) + threadIdx.x;
            uint32_t h[8];
#if LB_MERKLE_SUBTREE_LATENCY_FORM
            const uint4* pp = reinterpret_cast<const uint4*>(a.layers[d - 1] + (size_t)i * 16);
            hash_children_latency(h, __ldcg(pp), __ldcg(pp + 1), __ldcg(pp + 2), __ldcg(pp + 3));
#else
            hash_node<true>(h, a.layers[d - 1], nullptr, 0, i, one);
#endif
            uint4* o = reinterpret_cast<uint4*>(a.layers[d] + (size_t)i * 8);
            o[0] = make_uint4(h[0], h[1], h[2], h[3]);
            o[1] = make_uint4(h[4], h[5], h[6], h[7]);
        }
    }
}

cudaError_t merkle_commit_subtree(const MerkleSubtreeArgs& args, cudaStream_t stream) {
    if (args.depth < 1 || args.depth > MERKLE_SUBTREE_MAX_DEPTH || args.depth > args.log_top + 1 || args.n_cols < 0 ||
        args.n_cols > MERKLE_SMALL_COLS)
        return cudaErrorInvalidValue;
    uint32_t n_top = 1u << args.log_top;
    uint32_t S = n_top < 512u ? n_top : 512u;
    launch_k(merkle_subtree_kernel, n_top / S, 256, 0, stream, args, 1u);
    return cudaGetLastError();
}

// out[k*n_cols + c] = cols[c][idx[k]]
__global__ void gather_rows_kernel(uint32_t* out, const uint32_t* const* cols, int n_cols, const uint32_t* idx, int n_idx) {
    pdl_wait();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cols * n_idx) return;
    int k = t / n_cols, c = t % n_cols;
    out[t] = cols[c][idx[k]];
}

cudaError_t gather_rows(uint32_t* d_out, const uint32_t* const* d_cols, int n_cols, const uint32_t* d_idx, int n_idx,
                        cudaStream_t stream) {
    int total = n_cols * n_idx;
    if (total == 0) return cudaSuccess;
    launch_k(gather_rows_kernel, (total + 255) / 256, 256, 0, stream, d_out, d_cols, n_cols, d_idx, n_idx);
    return cudaGetLastError();
}

}  // namespace lb
