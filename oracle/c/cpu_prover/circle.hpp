// CPU prover (oracle; TEST INFRASTRUCTURE ONLY).
//
// Circle group over M31, cosets, canonic domains and the per-domain tables (points in storage order, CFFT twiddles).
// Restates stwo core/circle.rs, core/poly/circle/{canonic,domain}.rs, core/poly/line.rs and
// prover/backend/*/circle.rs::precompute_twiddles @0790eba (un-vendored).  Reference call site:
// /root/reference/crates/prover/src/prover.rs:38-42.
#pragma once
#include <omp.h>

#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "field.hpp"

namespace cpu {

struct Pt {
    uint32_t x, y;
};
static inline Pt pt_add(Pt p, Pt q) {
    return {m_sub(m_mul(p.x, q.x), m_mul(p.y, q.y)), m_add(m_mul(p.x, q.y), m_mul(p.y, q.x))};
}
static inline uint32_t double_x(uint32_t x) { return m_sub(m_mul(2, m_mul(x, x)), 1); }

constexpr int CIRCLE_LOG_ORDER = 31;
constexpr uint32_t IDX_MASK = 0x7fffffffu;  // indices live in Z / 2^31

// CirclePointIndex::to_point: idx * G, G = (2, 1268011823)
static inline Pt index_to_point(uint32_t idx) {
    static Pt pow2[31];
    static std::once_flag once;
    std::call_once(once, [] {
        Pt g{2, 1268011823u};
        for (int j = 0; j < 31; j++) {
            pow2[j] = g;
            g = pt_add(g, g);
        }
    });
    idx &= IDX_MASK;
    Pt r{1, 0};
    for (int j = 0; idx; j++, idx >>= 1)
        if (idx & 1) r = pt_add(r, pow2[j]);
    return r;
}
static inline uint32_t subgroup_gen(int log_size) { return 1u << (CIRCLE_LOG_ORDER - log_size); }

struct Coset {
    uint32_t initial;
    int log_size;
    uint32_t step() const { return subgroup_gen(log_size); }
    static Coset odds(int log) { return {subgroup_gen(log + 1), log}; }
    static Coset half_odds(int log) { return {subgroup_gen(log + 2), log}; }
    uint32_t index_at(uint64_t i) const { return (uint32_t)((initial + (uint64_t)step() * i) & IDX_MASK); }
    Pt at(uint64_t i) const { return index_to_point(index_at(i)); }
};

static inline uint32_t bit_reverse(uint32_t i, int bits) {
    if (bits == 0) return i;
    i = ((i >> 1) & 0x55555555u) | ((i & 0x55555555u) << 1);
    i = ((i >> 2) & 0x33333333u) | ((i & 0x33333333u) << 2);
    i = ((i >> 4) & 0x0F0F0F0Fu) | ((i & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(i) >> (32 - bits);
}

// QM31 point arithmetic (OODS point, mask points)
struct QPt {
    QM x, y;
};
static inline QPt qpt_add(QPt p, QPt q) { return {p.x * q.x - p.y * q.y, p.x * q.y + p.y * q.x}; }
static inline QM qdouble_x(QM x) { return x * x * 2u - qm(1); }

// coset_vanishing (stwo core/constraints.rs): x after log_size-1 doublings of p - initial + step/2
static inline uint32_t coset_vanishing(Coset c, Pt p) {
    Pt q = pt_add(pt_add(p, index_to_point((0u - c.initial) & IDX_MASK)), index_to_point(c.step() >> 1));
    uint32_t x = q.x;
    for (int i = 1; i < c.log_size; i++) x = double_x(x);
    return x;
}
static inline QM coset_vanishing(Coset c, QPt p) {
    Pt a = index_to_point((0u - c.initial) & IDX_MASK), b = index_to_point(c.step() >> 1);
    QPt q = qpt_add(qpt_add(p, QPt{qm(a.x), qm(a.y)}), QPt{qm(b.x), qm(b.y)});
    QM x = q.x;
    for (int i = 1; i < c.log_size; i++) x = qdouble_x(x);
    return x;
}

// in-place Montgomery batch inversion of n (non-zero) M31 values
static inline void batch_inverse(uint32_t* v, size_t n) {
    const size_t CH = 1 << 12;
#pragma omp parallel for schedule(static) if (n >= (1u << 16))
    for (size_t c0 = 0; c0 < n; c0 += CH) {
        size_t m = n - c0 < CH ? n - c0 : CH;
        uint32_t pre[CH];
        uint32_t acc = 1;
        for (size_t i = 0; i < m; i++) {
            pre[i] = acc;
            acc = m_mul(acc, v[c0 + i]);
        }
        uint32_t inv = m_inv(acc);
        for (size_t i = m; i-- > 0;) {
            uint32_t x = v[c0 + i];
            v[c0 + i] = m_mul(inv, pre[i]);
            inv = m_mul(inv, x);
        }
    }
}

// Tables of CanonicCoset(n).circle_domain() (half coset = half_odds(n - 1)):
//   hx[k], hy[k] (k < 2^(n-1)): the half-coset point at natural index bitrev(k); the domain point of storage row j is
//   (hx[j >> 1], j odd ? -hy[j >> 1] : hy[j >> 1]).
//   tw[0] = hy (circle layer), tw[i][h] (h < 2^(n-1-i)) = x of the (i-1)-times doubled half coset at bitrev(h): the line
//   layers; itw = their inverses.  tw[i+1][h] = double_x(tw[i][2h]).
struct Domain {
    int n;
    std::vector<uint32_t> hx;
    std::vector<std::vector<uint32_t>> tw, itw;
    const uint32_t* hy() const { return tw[0].data(); }
    uint32_t inv_n;
};

static inline void coset_points_natural(Coset c, std::vector<uint32_t>& xs, std::vector<uint32_t>& ys) {
    size_t n = (size_t)1 << c.log_size;
    xs.resize(n);
    ys.resize(n);
    Pt p0 = index_to_point(c.initial);
    xs[0] = p0.x;
    ys[0] = p0.y;
    for (int j = 0; j < c.log_size; j++) {
        Pt s = index_to_point((uint32_t)(((uint64_t)c.step() << j) & IDX_MASK));
        size_t h = (size_t)1 << j;
#pragma omp parallel for schedule(static) if (h >= (1u << 14))
        for (size_t i = 0; i < h; i++) {
            Pt q = pt_add(Pt{xs[i], ys[i]}, s);
            xs[h + i] = q.x;
            ys[h + i] = q.y;
        }
    }
}

static inline std::shared_ptr<const Domain> get_domain(int n) {
    static std::map<int, std::shared_ptr<const Domain>> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    auto d = std::make_shared<Domain>();
    d->n = n;
    d->inv_n = m_inv((uint32_t)(((uint64_t)1 << n) % P));
    if (n >= 1) {
        std::vector<uint32_t> xs, ys;
        coset_points_natural(Coset::half_odds(n - 1), xs, ys);
        size_t h = (size_t)1 << (n - 1);
        d->hx.resize(h);
        d->tw.resize(n);
        d->tw[0].resize(h);
#pragma omp parallel for schedule(static) if (h >= (1u << 14))
        for (size_t k = 0; k < h; k++) {
            uint32_t j = bit_reverse((uint32_t)k, n - 1);
            d->hx[k] = xs[j];
            d->tw[0][k] = ys[j];
        }
        for (int i = 1; i < n; i++) {
            size_t m = (size_t)1 << (n - 1 - i);
            d->tw[i].resize(m);
            const uint32_t* src = i == 1 ? d->hx.data() : d->tw[i - 1].data();
#pragma omp parallel for schedule(static) if (m >= (1u << 14))
            for (size_t k = 0; k < m; k++) d->tw[i][k] = i == 1 ? src[2 * k] : double_x(src[2 * k]);
        }
        d->itw = d->tw;
        for (int i = 0; i < n; i++) {
            d->itw[i].resize(d->tw[i].size() + W);  // padded: the packed low layers may read a full vector
            for (size_t k = d->tw[i].size(); k < d->itw[i].size(); k++) d->itw[i][k] = 1;
            batch_inverse(d->itw[i].data(), d->tw[i].size());
            d->tw[i].resize(d->tw[i].size() + W, 1);
        }
    }
    cache[n] = d;
    return d;
}

}  // namespace cpu
