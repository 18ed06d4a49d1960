// CPU prover (oracle; TEST INFRASTRUCTURE ONLY).
//
// Circle FFT over M31: interpolate / evaluate / eval_at_point.  Restates stwo prover/backend/cpu/circle.rs (the
// butterfly network, layer order and twiddle indexing) with the blocking idea of prover/backend/simd/{fft,circle}.rs
// @0790eba (packed butterflies, the layers below the vector width done by lane permutations, a cache-blocked schedule
// over the rayon pool - /root/reference/Cargo.toml:21-26 enables stwo's "parallel" feature).  Call sites in the
// reference: every tree_builder.extend_evals / commit, /root/reference/crates/prover/src/prover.rs:57,59,179,298.
//
//   evaluate   : layers i = n-1 .. 1 (twiddle x), then i = 0 (twiddle y); idx0 = (h << (i+1)) + l, idx1 = idx0 + 2^i;
//                (v0, v1) <- (v0 + v1 t, v0 - v1 t)
//   interpolate: i = 0 .. n-1 with (v0, v1) <- (v0 + v1, (v0 - v1) / t), then * 2^-n
#pragma once
#include <algorithm>

#include "circle.hpp"

namespace cpu {

constexpr int LOW_LOG = 12;    // layers [0, LOW_LOG) run on cache-resident blocks of 2^LOW_LOG elements
constexpr int GROUP_LOG = 6;   // the layers above in groups of <= GROUP_LOG, staged through a scratch tile
constexpr int TILE = 64;       // elements per row of that tile

static inline void bf_fwd(V& a, V& b, V t) {
    V tmp = b * t;
    b = a - tmp;
    a = a + tmp;
}
static inline void bf_inv(V& a, V& b, V t) {
    V s = a + b;
    b = (a - b) * t;
    a = s;
}

// scalar network for sizes below two vectors
static inline void cfft_scalar(uint32_t* v, int n, const Domain& d, bool fwd) {
    size_t N = (size_t)1 << n;
    if (fwd) {
        for (int i = n - 1; i >= 0; i--)
            for (size_t h = 0; h < (N >> (i + 1)); h++)
                for (size_t l = 0; l < ((size_t)1 << i); l++) {
                    size_t i0 = (h << (i + 1)) + l, i1 = i0 + ((size_t)1 << i);
                    uint32_t tmp = m_mul(v[i1], d.tw[i][h]);
                    uint32_t a = v[i0];
                    v[i0] = m_add(a, tmp);
                    v[i1] = m_sub(a, tmp);
                }
    } else {
        for (int i = 0; i < n; i++)
            for (size_t h = 0; h < (N >> (i + 1)); h++)
                for (size_t l = 0; l < ((size_t)1 << i); l++) {
                    size_t i0 = (h << (i + 1)) + l, i1 = i0 + ((size_t)1 << i);
                    uint32_t a = v[i0], b = v[i1];
                    v[i0] = m_add(a, b);
                    v[i1] = m_mul(m_sub(a, b), d.itw[i][h]);
                }
        for (size_t k = 0; k < N; k++) v[k] = m_mul(v[k], d.inv_n);
    }
}

// layers 0 .. LOGW of the 2W elements at p (element index base): each vdeinterleave moves the next index bit into the
// "which register" position, so every layer is a butterfly between the two registers; LOGW+1 of them restore the layout.
template <int I>
static inline void low_inv_step(V& a, V& b, const Domain& d, size_t base) {
    V lo, hi;
    vdeinterleave(a, b, lo, hi);
    V t = vload_pattern<(W >> I)>(d.itw[I].data() + (base >> (I + 1)));
    bf_inv(lo, hi, t);
    a = lo;
    b = hi;
}
template <int I>
static inline void low_fwd_step(V& a, V& b, const Domain& d, size_t base) {
    V t = vload_pattern<(W >> I)>(d.tw[I].data() + (base >> (I + 1)));
    bf_fwd(a, b, t);
    V x, y;
    vinterleave(a, b, x, y);
    a = x;
    b = y;
}

// all layers below lowlog (<= LOW_LOG) of the block of 2^lowlog elements starting at element index `base`
static inline void low_block(uint32_t* v, size_t base, int lowlog, const Domain& d, bool fwd) {
    size_t B = (size_t)1 << lowlog;
    uint32_t* p = v + base;
    if (!fwd) {
        for (size_t g = 0; g < B; g += 2 * W) {
            V a = vload(p + g), b = vload(p + g + W);
            low_inv_step<0>(a, b, d, base + g);
            low_inv_step<1>(a, b, d, base + g);
            low_inv_step<2>(a, b, d, base + g);
            low_inv_step<3>(a, b, d, base + g);
#if CPU_LOGW == 4
            low_inv_step<4>(a, b, d, base + g);
#endif
            vstore(p + g, a);
            vstore(p + g + W, b);
        }
        for (int i = LOGW + 1; i < lowlog; i++) {
            size_t s = (size_t)1 << i;
            for (size_t h0 = 0; h0 < B; h0 += 2 * s) {
                V t = vset1(d.itw[i][(base + h0) >> (i + 1)]);
                for (size_t l = 0; l < s; l += W) {
                    V a = vload(p + h0 + l), b = vload(p + h0 + l + s);
                    bf_inv(a, b, t);
                    vstore(p + h0 + l, a);
                    vstore(p + h0 + l + s, b);
                }
            }
        }
    } else {
        for (int i = lowlog - 1; i > LOGW; i--) {
            size_t s = (size_t)1 << i;
            for (size_t h0 = 0; h0 < B; h0 += 2 * s) {
                V t = vset1(d.tw[i][(base + h0) >> (i + 1)]);
                for (size_t l = 0; l < s; l += W) {
                    V a = vload(p + h0 + l), b = vload(p + h0 + l + s);
                    bf_fwd(a, b, t);
                    vstore(p + h0 + l, a);
                    vstore(p + h0 + l + s, b);
                }
            }
        }
        for (size_t g = 0; g < B; g += 2 * W) {
            V a = vload(p + g), b = vload(p + g + W);
#if CPU_LOGW == 4
            low_fwd_step<4>(a, b, d, base + g);
#endif
            low_fwd_step<3>(a, b, d, base + g);
            low_fwd_step<2>(a, b, d, base + g);
            low_fwd_step<1>(a, b, d, base + g);
            low_fwd_step<0>(a, b, d, base + g);
            vstore(p + g, a);
            vstore(p + g + W, b);
        }
    }
}

// layers [s, e) of the rows {hi * 2^e + r * 2^s + lo0 + [0, TILE)}, r < 2^(e-s), through a scratch tile
static inline void high_tile(uint32_t* v, size_t hi, size_t lo0, int s, int e, const Domain& d, bool fwd, uint32_t scale) {
    const int R = 1 << (e - s);
    alignas(64) uint32_t t[(1 << GROUP_LOG) * TILE];
    uint32_t* p = v + (hi << e) + lo0;
    for (int r = 0; r < R; r++) memcpy(t + r * TILE, p + ((size_t)r << s), TILE * 4);
    auto layer = [&](int j) {
        int dd = 1 << (j - s);
        const uint32_t* tw = fwd ? d.tw[j].data() : d.itw[j].data();
        for (int rb = 0; rb < R; rb += 2 * dd)
            for (int r = rb; r < rb + dd; r++) {
                V tv = vset1(tw[(hi << (e - j - 1)) + (r >> (j - s + 1))]);
                uint32_t *x = t + r * TILE, *y = t + (r + dd) * TILE;
                for (int k = 0; k < TILE; k += W) {
                    V a = vload(x + k), b = vload(y + k);
                    if (fwd)
                        bf_fwd(a, b, tv);
                    else
                        bf_inv(a, b, tv);
                    vstore(x + k, a);
                    vstore(y + k, b);
                }
            }
    };
    if (fwd)
        for (int j = e - 1; j >= s; j--) layer(j);
    else
        for (int j = s; j < e; j++) layer(j);
    if (scale != 1) {
        V sc = vset1(scale);
        for (int k = 0; k < R * TILE; k += W) vstore(t + k, vload(t + k) * sc);
    }
    for (int r = 0; r < R; r++) memcpy(p + ((size_t)r << s), t + r * TILE, TILE * 4);
}

// In-place transform of ncols columns of 2^n values each.  fwd: coefficients -> evaluations on CanonicCoset(n)'s circle
// domain; !fwd: evaluations -> coefficients (scaled by 2^-n).
static inline void cfft(uint32_t* const* cols, int ncols, int n, bool fwd) {
    if (n == 0) return;
    auto dom = get_domain(n);
    const Domain& d = *dom;
    if (n < LOGW + 1) {
        for (int c = 0; c < ncols; c++) cfft_scalar(cols[c], n, d, fwd);
        return;
    }
    const int lowlog = std::min(n, LOW_LOG);
    // split the layers above lowlog into groups of <= GROUP_LOG
    std::vector<std::pair<int, int>> groups;
    {
        int rem = n - lowlog, ng = (rem + GROUP_LOG - 1) / GROUP_LOG, s = lowlog;
        for (int g = 0; g < ng; g++) {
            int len = rem / ng + (g < rem % ng ? 1 : 0);
            groups.push_back({s, s + len});
            s += len;
        }
    }
    const size_t N = (size_t)1 << n;
    auto run_low = [&](uint32_t scale) {
        size_t nb = N >> lowlog, items = nb * ncols;
#pragma omp parallel for schedule(static) if (items > 1 && N * ncols >= (1u << 14))
        for (size_t it = 0; it < items; it++) {
            uint32_t* v = cols[it / nb];
            size_t base = (it % nb) << lowlog;
            low_block(v, base, lowlog, d, fwd);
            if (scale != 1) {
                V sc = vset1(scale);
                for (size_t k = 0; k < ((size_t)1 << lowlog); k += W) vstore(v + base + k, vload(v + base + k) * sc);
            }
        }
    };
    auto run_group = [&](std::pair<int, int> g, uint32_t scale) {
        int s = g.first, e = g.second;
        size_t nhi = N >> e, ntile = ((size_t)1 << s) / TILE, per = nhi * ntile, items = per * ncols;
#pragma omp parallel for schedule(static)
        for (size_t it = 0; it < items; it++) {
            uint32_t* v = cols[it / per];
            size_t w = it % per;
            high_tile(v, w / ntile, (w % ntile) * TILE, s, e, d, fwd, scale);
        }
    };
    if (!fwd) {
        run_low(groups.empty() ? d.inv_n : 1);
        for (size_t g = 0; g < groups.size(); g++) run_group(groups[g], g + 1 == groups.size() ? d.inv_n : 1);
    } else {
        for (size_t g = groups.size(); g-- > 0;) run_group(groups[g], 1);
        run_low(1);
    }
}

// PolyOps::eval_at_point: sum_j c_j * prod_{bits b of j} m_b with m = [y, x, pi(x), pi^2(x), ...] (stwo's fold
// computes the same sum; field arithmetic is exact, so the association order does not matter).
// Evaluates ncols polynomials of 2^n coefficients at one point.
static inline void eval_at_point(const uint32_t* const* cols, int ncols, int n, QPt pt, QM* out) {
    if (n == 0) {
        for (int c = 0; c < ncols; c++) out[c] = qm(cols[c][0]);
        return;
    }
    std::vector<QM> m(n);
    m[0] = pt.y;
    QM x = pt.x;
    for (int i = 1; i < n; i++) {
        m[i] = x;
        x = qdouble_x(x);
    }
    const int lb = std::min(n, 10);
    const size_t LB = (size_t)1 << lb, NB = (size_t)1 << (n - lb);
    // low table in coordinate planes, high table as scalars
    std::vector<uint32_t> lowt(4 * LB);
    std::vector<QM> low(LB), high(NB);
    low[0] = qm(1);
    for (int b = 0; b < lb; b++)
        for (size_t j = 0; j < ((size_t)1 << b); j++) low[j + ((size_t)1 << b)] = low[j] * m[b];
    high[0] = qm(1);
    for (int b = 0; b < n - lb; b++)
        for (size_t j = 0; j < ((size_t)1 << b); j++) high[j + ((size_t)1 << b)] = high[j] * m[lb + b];
    for (size_t j = 0; j < LB; j++)
        for (int k = 0; k < 4; k++) lowt[k * LB + j] = low[j].c[k];
    size_t items = (size_t)ncols * NB;
    std::vector<QM> partial(items);
#pragma omp parallel for schedule(static) if (items * LB >= (1u << 14))
    for (size_t it = 0; it < items; it++) {
        const uint32_t* c = cols[it / NB] + ((it % NB) << lb);
        QM acc = qm(0);
        if (LB >= (size_t)W) {
            V a[4] = {vzero(), vzero(), vzero(), vzero()};
            for (size_t j = 0; j < LB; j += W) {
                V cv = vload(c + j);
                for (int k = 0; k < 4; k++) a[k] = a[k] + cv * vload(lowt.data() + k * LB + j);
            }
            for (int k = 0; k < 4; k++) {
                uint32_t t[W];
                vstore(t, a[k]);
                uint32_t s = 0;
                for (int l = 0; l < W; l++) s = m_add(s, t[l]);
                acc.c[k] = s;
            }
        } else {
            for (size_t j = 0; j < LB; j++) acc = acc + low[j] * c[j];
        }
        partial[it] = acc * high[it % NB];
    }
    for (int c = 0; c < ncols; c++) {
        QM s = qm(0);
        for (size_t b = 0; b < NB; b++) s = s + partial[(size_t)c * NB + b];
        out[c] = s;
    }
}

}  // namespace cpu
