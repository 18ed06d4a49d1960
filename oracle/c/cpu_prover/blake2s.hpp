// CPU prover (oracle; TEST INFRASTRUCTURE ONLY).
//
// Blake2s-256 (RFC 7693): a scalar hash / raw compression for the Fiat-Shamir channel and grind, and a packed
// compression (one hash per vector lane) for the Merkle layers - the shape of stwo prover/backend/simd/blake2s.rs.
// Channel: stwo core/channel/blake2s.rs; hasher: core/vcs/blake2_hash.rs, blake2_merkle.rs @0790eba (un-vendored;
// /root/reference/Cargo.lock:198-199 pins blake2 0.10.6).  Reference call sites: crates/prover/src/prover.rs:44,59,177,179.
#pragma once
#include <array>
#include <string>
#include <vector>

#include "field.hpp"

namespace cpu {

static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

// F(h, m, t, f): h[8] updated in place
static inline void b2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1, uint32_t f0, uint32_t f1) {
    uint32_t v[16];
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[i + 8] = B2S_IV[i];
    }
    v[12] ^= t0;
    v[13] ^= t1;
    v[14] ^= f0;
    v[15] ^= f1;
#define CPU_G(a, b, c, d, x, y)             \
    v[a] = v[a] + v[b] + (x);               \
    v[d] = rotr32(v[d] ^ v[a], 16);         \
    v[c] = v[c] + v[d];                     \
    v[b] = rotr32(v[b] ^ v[c], 12);         \
    v[a] = v[a] + v[b] + (y);               \
    v[d] = rotr32(v[d] ^ v[a], 8);          \
    v[c] = v[c] + v[d];                     \
    v[b] = rotr32(v[b] ^ v[c], 7);
    for (int r = 0; r < 10; r++) {
        const uint8_t* s = B2S_SIGMA[r];
        CPU_G(0, 4, 8, 12, m[s[0]], m[s[1]])
        CPU_G(1, 5, 9, 13, m[s[2]], m[s[3]])
        CPU_G(2, 6, 10, 14, m[s[4]], m[s[5]])
        CPU_G(3, 7, 11, 15, m[s[6]], m[s[7]])
        CPU_G(0, 5, 10, 15, m[s[8]], m[s[9]])
        CPU_G(1, 6, 11, 12, m[s[10]], m[s[11]])
        CPU_G(2, 7, 8, 13, m[s[12]], m[s[13]])
        CPU_G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef CPU_G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

using Hash = std::array<uint8_t, 32>;

static inline Hash b2s_hash(const uint8_t* data, size_t n) {
    uint32_t h[8];
    for (int i = 0; i < 8; i++) h[i] = B2S_IV[i];
    h[0] ^= 0x01010020u;
    size_t off = 0;
    uint64_t t = 0;
    do {
        size_t len = n - off < 64 ? n - off : 64;
        bool last = off + len == n;
        uint8_t blk[64] = {0};
        if (len) memcpy(blk, data + off, len);
        uint32_t m[16];
        memcpy(m, blk, 64);
        t += len;
        b2s_compress(h, m, (uint32_t)t, (uint32_t)(t >> 32), last ? 0xFFFFFFFFu : 0, 0);
        off += len;
    } while (off < n);
    Hash out;
    memcpy(out.data(), h, 32);
    return out;
}
static inline Hash b2s_hash(const std::vector<uint8_t>& v) { return b2s_hash(v.data(), v.size()); }

// packed compression: lane l of h[i] / m[i] belongs to hash l
static inline void b2s_compress_v(V h[8], const V m[16], uint32_t t0, uint32_t f0) {
    V v[16];
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[i + 8] = vset1(B2S_IV[i]);
    }
    v[12] = vxor(v[12], vset1(t0));
    v[14] = vxor(v[14], vset1(f0));
#define CPU_GV(a, b, c, d, x, y)                    \
    v[a] = vadd32(vadd32(v[a], v[b]), (x));         \
    v[d] = vror<16>(vxor(v[d], v[a]));              \
    v[c] = vadd32(v[c], v[d]);                      \
    v[b] = vror<12>(vxor(v[b], v[c]));              \
    v[a] = vadd32(vadd32(v[a], v[b]), (y));         \
    v[d] = vror<8>(vxor(v[d], v[a]));               \
    v[c] = vadd32(v[c], v[d]);                      \
    v[b] = vror<7>(vxor(v[b], v[c]));
    for (int r = 0; r < 10; r++) {
        const uint8_t* s = B2S_SIGMA[r];
        CPU_GV(0, 4, 8, 12, m[s[0]], m[s[1]])
        CPU_GV(1, 5, 9, 13, m[s[2]], m[s[3]])
        CPU_GV(2, 6, 10, 14, m[s[4]], m[s[5]])
        CPU_GV(3, 7, 11, 15, m[s[6]], m[s[7]])
        CPU_GV(0, 5, 10, 15, m[s[8]], m[s[9]])
        CPU_GV(1, 6, 11, 12, m[s[10]], m[s[11]])
        CPU_GV(2, 7, 8, 13, m[s[12]], m[s[13]])
        CPU_GV(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef CPU_GV
    for (int i = 0; i < 8; i++) h[i] = vxor(h[i], vxor(v[i], v[i + 8]));
}

// ---- Blake2sChannel ("legacy" = the variant pinned by the reference's committed proof ui/demo/public/proof; "v2" =
// the later mixing) ---------------------------------------------------------------------------------
struct Channel {
    int variant = 0;  // 0 legacy, 1 v2
    Hash digest{};
    uint32_t n_sent = 0;

    void update(const Hash& d) {
        digest = d;
        n_sent = 0;
    }
    void mix_bytes(const uint8_t* p, size_t n) {
        std::vector<uint8_t> buf(32 + n);
        memcpy(buf.data(), digest.data(), 32);
        if (n) memcpy(buf.data() + 32, p, n);
        update(b2s_hash(buf));
    }
    void mix_root(const Hash& root) { mix_bytes(root.data(), 32); }
    void mix_felts(const QM* f, size_t n) { mix_bytes((const uint8_t*)f, 16 * n); }
    void mix_u64(uint64_t v) {
        uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        if (variant == 0) {
            uint32_t h[8], m[16] = {lo, hi};
            memcpy(h, digest.data(), 32);
            b2s_compress(h, m, 0, 0, 0, 0);
            Hash d;
            memcpy(d.data(), h, 32);
            update(d);
        } else {
            uint32_t w[2] = {lo, hi};
            mix_bytes((const uint8_t*)w, 8);
        }
    }
    Hash draw_random_bytes() {
        uint8_t buf[65] = {0};
        memcpy(buf, digest.data(), 32);
        memcpy(buf + 32, &n_sent, 4);
        n_sent++;
        return b2s_hash(buf, variant == 1 ? 65 : 64);
    }
    void draw_base_felts(uint32_t out[8]) {
        for (;;) {
            Hash r = draw_random_bytes();
            uint32_t u[8];
            memcpy(u, r.data(), 32);
            bool ok = true;
            for (int i = 0; i < 8; i++) ok &= u[i] < 2 * P;
            if (!ok) continue;
            for (int i = 0; i < 8; i++) out[i] = u[i] >= P ? u[i] - P : u[i];
            return;
        }
    }
    QM draw_secure_felt() {
        uint32_t f[8];
        draw_base_felts(f);
        return qm(f[0], f[1], f[2], f[3]);
    }
    void draw_secure_felts(QM* out, int n) {
        uint32_t pool[8];
        int have = 0, pos = 0;
        for (int k = 0; k < n; k++) {
            if (have - pos < 4) {
                draw_base_felts(pool);  // n is even at every call site: the pool is always drained
                have = 8;
                pos = 0;
            }
            out[k] = qm(pool[pos], pool[pos + 1], pool[pos + 2], pool[pos + 3]);
            pos += 4;
        }
    }
    int trailing_zeros() const {
        for (int i = 0; i < 16; i++)
            if (digest[i]) return 8 * i + __builtin_ctz(digest[i]);
        return 128;
    }
};

// GrindOps<Blake2sChannel>::grind: smallest nonce whose mix_u64 leaves >= pow_bits trailing zeros
static inline uint64_t grind(const Channel& ch, uint32_t pow_bits) {
    for (uint64_t nonce = 0;; nonce++) {
        Channel c = ch;
        c.mix_u64(nonce);
        if (ch.variant == 0) {
            uint32_t w0;
            memcpy(&w0, c.digest.data(), 4);
            uint32_t mask = pow_bits >= 32 ? 0xFFFFFFFFu : ((1u << pow_bits) - 1);
            if ((w0 & mask) == 0) return nonce;
        } else if ((uint32_t)c.trailing_zeros() >= pow_bits) {
            return nonce;
        }
    }
}

}  // namespace cpu
