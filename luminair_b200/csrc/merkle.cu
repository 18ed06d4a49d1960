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
#include "merkle.cuh"

namespace lb {

__device__ __constant__ uint32_t c_iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                            0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

__device__ __forceinline__ uint32_t rotr16(uint32_t x) { return __byte_perm(x, x, 0x1032); }
__device__ __forceinline__ uint32_t rotr8(uint32_t x) { return __byte_perm(x, x, 0x0321); }
__device__ __forceinline__ uint32_t rotr12(uint32_t x) { return __funnelshift_r(x, x, 12); }
__device__ __forceinline__ uint32_t rotr7(uint32_t x) { return __funnelshift_r(x, x, 7); }

#define B2S_G(a, b, c, d, x, y) \
    a = a + b + (x);            \
    d = rotr16(d ^ a);          \
    c = c + d;                  \
    b = rotr12(b ^ c);          \
    a = a + b + (y);            \
    d = rotr8(d ^ a);           \
    c = c + d;                  \
    b = rotr7(b ^ c);

#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    B2S_G(v0, v4, v8, v12, m[s0], m[s1])                                                 \
    B2S_G(v1, v5, v9, v13, m[s2], m[s3])                                                 \
    B2S_G(v2, v6, v10, v14, m[s4], m[s5])                                                \
    B2S_G(v3, v7, v11, v15, m[s6], m[s7])                                                \
    B2S_G(v0, v5, v10, v15, m[s8], m[s9])                                                \
    B2S_G(v1, v6, v11, v12, m[s10], m[s11])                                              \
    B2S_G(v2, v7, v8, v13, m[s12], m[s13])                                               \
    B2S_G(v3, v4, v9, v14, m[s14], m[s15])

__device__ __forceinline__ void blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1,
                                                 uint32_t f0) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
    uint32_t v12 = 0x510E527Fu ^ t0, v13 = 0x9B05688Cu ^ t1, v14 = 0x1F83D9ABu ^ f0, v15 = 0x5BE0CD19u;
    B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8;
    h[1] ^= v1 ^ v9;
    h[2] ^= v2 ^ v10;
    h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12;
    h[5] ^= v5 ^ v13;
    h[6] ^= v6 ^ v14;
    h[7] ^= v7 ^ v15;
}

__device__ __forceinline__ void blake2s_init(uint32_t h[8]) {
    h[0] = 0x6A09E667u ^ 0x01010020u;
    h[1] = 0xBB67AE85u;
    h[2] = 0x3C6EF372u;
    h[3] = 0xA54FF53Au;
    h[4] = 0x510E527Fu;
    h[5] = 0x9B05688Cu;
    h[6] = 0x1F83D9ABu;
    h[7] = 0x5BE0CD19u;
}

// cols: device array of n_cols column base pointers (each 2^log_size u32).
__global__ void __launch_bounds__(256) merkle_layer_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ prev,
                                                           const uint32_t* const* __restrict__ cols, int n_cols,
                                                           uint32_t n_nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t h[8];
    blake2s_init(h);
    uint32_t m[16];
    uint32_t t = 0;
    int c = 0;
    bool have_block = false;
    if (prev) {
        const uint4* pp = reinterpret_cast<const uint4*>(prev + (size_t)i * 16);
        uint4 a = pp[0], b = pp[1], cc = pp[2], d = pp[3];
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
            blake2s_compress(h, m, t, 0, last ? 0xFFFFFFFFu : 0u);
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
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 8);
    o[0] = make_uint4(h[0], h[1], h[2], h[3]);
    o[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

cudaError_t merkle_commit_layer(uint32_t* out, const uint32_t* prev, const uint32_t* const* d_cols, int n_cols,
                                int log_size, cudaStream_t stream) {
    uint32_t n = 1u << log_size;
    merkle_layer_kernel<<<(n + 255) / 256, 256, 0, stream>>>(out, prev, d_cols, n_cols, n);
    return cudaGetLastError();
}

// out[k*n_cols + c] = cols[c][idx[k]]
__global__ void gather_rows_kernel(uint32_t* out, const uint32_t* const* cols, int n_cols, const uint32_t* idx, int n_idx) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cols * n_idx) return;
    int k = t / n_cols, c = t % n_cols;
    out[t] = cols[c][idx[k]];
}

cudaError_t gather_rows(uint32_t* d_out, const uint32_t* const* d_cols, int n_cols, const uint32_t* d_idx, int n_idx,
                        cudaStream_t stream) {
    int total = n_cols * n_idx;
    if (total == 0) return cudaSuccess;
    gather_rows_kernel<<<(total + 255) / 256, 256, 0, stream>>>(d_out, d_cols, n_cols, d_idx, n_idx);
    return cudaGetLastError();
}

}  // namespace lb
