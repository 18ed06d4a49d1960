// Blake2s-256 compression shared by the device kernels (Merkle layers, grind) and the host side
// of the prover (Fiat-Shamir channel, empty-tree root).  Follows RFC 7693; the way stwo uses it
// (core/vcs/blake2_hash.rs, core/channel/blake2s.rs @0790eba, un-vendored dependency of
// /root/reference, Cargo.toml:21-28) is restated where it is called.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

namespace lb {

#ifdef __CUDA_ARCH__
__device__ __forceinline__ uint32_t b2_rotr16(uint32_t x) { return __byte_perm(x, x, 0x1032); }
__device__ __forceinline__ uint32_t b2_rotr8(uint32_t x) { return __byte_perm(x, x, 0x0321); }
__device__ __forceinline__ uint32_t b2_rotr12(uint32_t x) { return __funnelshift_r(x, x, 12); }
__device__ __forceinline__ uint32_t b2_rotr7(uint32_t x) { return __funnelshift_r(x, x, 7); }
#else
inline uint32_t b2_rotr16(uint32_t x) { return (x >> 16) | (x << 16); }
inline uint32_t b2_rotr8(uint32_t x) { return (x >> 8) | (x << 24); }
inline uint32_t b2_rotr12(uint32_t x) { return (x >> 12) | (x << 20); }
inline uint32_t b2_rotr7(uint32_t x) { return (x >> 7) | (x << 25); }
#endif

// LB_B2S_ADD3 / LB_B2S_ADD2 are defined by the function that expands the rounds: plain adds on the host
// and in the generic device version; IMAD-forced adds in blake2s_compress_fma (see below).
#define LB_B2S_G(a, b, c, d, x, y) \
    a = LB_B2S_ADD3(a, b, (x));    \
    d = b2_rotr16(d ^ a);          \
    c = LB_B2S_ADD2(c, d);         \
    b = b2_rotr12(b ^ c);          \
    a = LB_B2S_ADD3(a, b, (y));    \
    d = b2_rotr8(d ^ a);           \
    c = LB_B2S_ADD2(c, d);         \
    b = b2_rotr7(b ^ c);

#define LB_B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    LB_B2S_G(v0, v4, v8, v12, m[s0], m[s1])                                                 \
    LB_B2S_G(v1, v5, v9, v13, m[s2], m[s3])                                                 \
    LB_B2S_G(v2, v6, v10, v14, m[s4], m[s5])                                                \
    LB_B2S_G(v3, v7, v11, v15, m[s6], m[s7])                                                \
    LB_B2S_G(v0, v5, v10, v15, m[s8], m[s9])                                                \
    LB_B2S_G(v1, v6, v11, v12, m[s10], m[s11])                                              \
    LB_B2S_G(v2, v7, v8, v13, m[s12], m[s13])                                               \
    LB_B2S_G(v3, v4, v9, v14, m[s14], m[s15])

#define LB_B2S_ADD3(p, q, r) ((p) + (q) + (r))
#define LB_B2S_ADD2(p, q) ((p) + (q))
// h <- compress(h, m, t = (t0, t1), f0)
__host__ __device__ __forceinline__ void blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1,
                                                          uint32_t f0) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
    uint32_t v12 = 0x510E527Fu ^ t0, v13 = 0x9B05688Cu ^ t1, v14 = 0x1F83D9ABu ^ f0, v15 = 0x5BE0CD19u;
    LB_B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    LB_B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    LB_B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    LB_B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    LB_B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    LB_B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    LB_B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    LB_B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    LB_B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    LB_B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8;
    h[1] ^= v1 ^ v9;
    h[2] ^= v2 ^ v10;
    h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12;
    h[5] ^= v5 ^ v13;
    h[6] ^= v6 ^ v14;
    h[7] ^= v7 ^ v15;
}

#undef LB_B2S_ADD3
#undef LB_B2S_ADD2

#ifdef __CUDACC__
// Device variant for the throughput kernels.  Blake2s is 320 adds + 320 xors + 320 rotates per block; xor
// and rotate only run on the ALU pipe (LOP3 / SHF / PRMT, one warp-instruction per 2 clk per sub-partition), so
// every add is forced onto the FMA pipe as IMAD x*1+y (`one` must be an opaque kernel parameter equal to 1,
// otherwise ptxas folds it back into IADD3): ALU 640 instructions, FMA 480, instead of ALU 800 / FMA 160.
__device__ __forceinline__ uint32_t b2_mad1(uint32_t p, uint32_t one, uint32_t q) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(p), "r"(one), "r"(q));
    return r;
}
#define LB_B2S_ADD3(p, q, r) b2_mad1(b2_mad1((p), one, (q)), one, (r))
#define LB_B2S_ADD2(p, q) b2_mad1((p), one, (q))
__device__ __forceinline__ void blake2s_compress_fma(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1,
                                                      uint32_t f0, uint32_t one) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
    uint32_t v12 = 0x510E527Fu ^ t0, v13 = 0x9B05688Cu ^ t1, v14 = 0x1F83D9ABu ^ f0, v15 = 0x5BE0CD19u;
    LB_B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    LB_B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    LB_B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    LB_B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    LB_B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    LB_B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    LB_B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    LB_B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    LB_B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    LB_B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8;
    h[1] ^= v1 ^ v9;
    h[2] ^= v2 ^ v10;
    h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12;
    h[5] ^= v5 ^ v13;
    h[6] ^= v6 ^ v14;
    h[7] ^= v7 ^ v15;
}

#undef LB_B2S_ADD3
#undef LB_B2S_ADD2
#endif

__host__ __device__ __forceinline__ void blake2s_init(uint32_t h[8]) {
    h[0] = 0x6A09E667u ^ 0x01010020u;  // digest length 32, no key, fanout = depth = 1
    h[1] = 0xBB67AE85u;
    h[2] = 0x3C6EF372u;
    h[3] = 0xA54FF53Au;
    h[4] = 0x510E527Fu;
    h[5] = 0x9B05688Cu;
    h[6] = 0x1F83D9ABu;
    h[7] = 0x5BE0CD19u;
}

// ---- host-only byte interface (little-endian host) ------------------------------------------
struct Hash32 {
    uint8_t b[32];
    bool operator==(const Hash32& o) const { return std::memcmp(b, o.b, 32) == 0; }
};

inline Hash32 blake2s_hash(const uint8_t* data, size_t len) {
    uint32_t h[8];
    blake2s_init(h);
    uint32_t m[16];
    size_t off = 0;
    while (len - off > 64) {
        std::memcpy(m, data + off, 64);
        off += 64;
        blake2s_compress(h, m, (uint32_t)off, (uint32_t)((uint64_t)off >> 32), 0);
    }
    std::memset(m, 0, 64);
    std::memcpy(m, data + off, len - off);
    blake2s_compress(h, m, (uint32_t)len, (uint32_t)((uint64_t)len >> 32), 0xFFFFFFFFu);
    Hash32 out;
    std::memcpy(out.b, h, 32);
    return out;
}

}  // namespace lb
