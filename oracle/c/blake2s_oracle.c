/* Blake2s-256 Merkle layer on the CPU (oracle / cpu_baseline; TEST INFRASTRUCTURE ONLY).
 * hash_node(children, column values) as in stwo core/vcs/blake2_merkle.rs (see oracle/merkle.py).
 * RFC 7693 Blake2s; OpenMP over nodes. */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
#define G(a, b, c, d, x, y)          \
    v[a] = v[a] + v[b] + (x);        \
    v[d] = rotr(v[d] ^ v[a], 16);    \
    v[c] = v[c] + v[d];              \
    v[b] = rotr(v[b] ^ v[c], 12);    \
    v[a] = v[a] + v[b] + (y);        \
    v[d] = rotr(v[d] ^ v[a], 8);     \
    v[c] = v[c] + v[d];              \
    v[b] = rotr(v[b] ^ v[c], 7);

void oracle_blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1, uint32_t f0, uint32_t f1) {
    uint32_t v[16];
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= t0; v[13] ^= t1; v[14] ^= f0; v[15] ^= f1;
    for (int r = 0; r < 10; ++r) {
        const uint8_t* s = SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]) G(1, 5, 9, 13, m[s[2]], m[s[3]])
        G(2, 6, 10, 14, m[s[4]], m[s[5]]) G(3, 7, 11, 15, m[s[6]], m[s[7]])
        G(0, 5, 10, 15, m[s[8]], m[s[9]]) G(1, 6, 11, 12, m[s[10]], m[s[11]])
        G(2, 7, 8, 13, m[s[12]], m[s[13]]) G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

/* out[i] = Blake2s(prev[2i] || prev[2i+1] || cols[0][i] .. cols[n_cols-1][i]); prev may be NULL */
void oracle_merkle_layer(uint32_t* out, const uint32_t* prev, const uint32_t* const* cols, int n_cols, size_t n_nodes) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_nodes; ++i) {
        uint32_t h[8], m[16];
        memcpy(h, IV, sizeof(h));
        h[0] ^= 0x01010020u;
        uint32_t t = 0;
        int c = 0, have = 0;
        if (prev) { memcpy(m, prev + i * 16, 64); t = 64; have = 1; }
        for (;;) {
            if (have) {
                int last = c >= n_cols;
                oracle_blake2s_compress(h, m, t, 0, last ? 0xFFFFFFFFu : 0, 0);
                if (last) break;
            }
            int take = n_cols - c < 16 ? n_cols - c : 16;
            memset(m, 0, sizeof(m));
            if (take <= 0) { have = 1; c = n_cols; continue; }
            for (int j = 0; j < take; ++j) m[j] = cols[c + j][i];
            c += take; t += 4u * take; have = 1;
        }
        memcpy(out + i * 8, h, 32);
    }
}
