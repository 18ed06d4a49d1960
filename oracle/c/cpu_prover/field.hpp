// CPU prover (oracle; TEST INFRASTRUCTURE ONLY - never linked into or called by luminair_b200).
//
// M31 / CM31 / QM31 arithmetic, scalar and packed.  Restates stwo core/fields/{m31,cm31,qm31}.rs and
// prover/backend/simd/{m31,cm31,qm31}.rs @0790eba (un-vendored; Cargo.lock:2055-2057): `PackedM31` is 16 u32 lanes
// there (N_LANES = 16, relied on by /root/reference/crates/air/src/utils.rs:22-27); here the lane count is the
// native vector width (16 with AVX-512, 8 with AVX2).
#pragma once
#include <immintrin.h>

#include <cstdint>
#include <cstring>

namespace cpu {

constexpr uint32_t P = 0x7fffffffu;

// ---- scalar M31 ---------------------------------------------------------------------------------
static inline uint32_t m_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return s >= P ? s - P : s;
}
static inline uint32_t m_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + P - b; }
static inline uint32_t m_neg(uint32_t a) { return a ? P - a : 0; }
static inline uint32_t m_reduce(uint64_t x) {  // x < 2^62
    uint64_t r = (x & P) + (x >> 31);
    r = (r & P) + (r >> 31);
    return r >= P ? (uint32_t)(r - P) : (uint32_t)r;
}
static inline uint32_t m_mul(uint32_t a, uint32_t b) { return m_reduce((uint64_t)a * b); }
static inline uint32_t m_pow(uint32_t a, uint64_t e) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = m_mul(r, a);
        a = m_mul(a, a);
        e >>= 1;
    }
    return r;
}
static inline uint32_t m_inv(uint32_t a) { return m_pow(a, P - 2); }

struct CM {
    uint32_t a, b;
};
static inline CM operator+(CM x, CM y) { return {m_add(x.a, y.a), m_add(x.b, y.b)}; }
static inline CM operator-(CM x, CM y) { return {m_sub(x.a, y.a), m_sub(x.b, y.b)}; }
static inline CM operator*(CM x, CM y) {
    return {m_sub(m_mul(x.a, y.a), m_mul(x.b, y.b)), m_add(m_mul(x.a, y.b), m_mul(x.b, y.a))};
}
static inline CM cm_inv(CM x) {
    uint32_t n = m_inv(m_add(m_mul(x.a, x.a), m_mul(x.b, x.b)));
    return {m_mul(x.a, n), m_mul(m_neg(x.b), n)};
}

// QM31 = CM31[u]/(u^2 - (2 + i)); coordinates (a.re, a.im, b.re, b.im)
struct QM {
    uint32_t c[4];
    CM lo() const { return {c[0], c[1]}; }
    CM hi() const { return {c[2], c[3]}; }
    bool operator==(const QM& o) const { return !memcmp(c, o.c, 16); }
    bool operator!=(const QM& o) const { return !(*this == o); }
    bool is_zero() const { return !(c[0] | c[1] | c[2] | c[3]); }
};
static inline QM qm(uint32_t a, uint32_t b = 0, uint32_t c = 0, uint32_t d = 0) { return QM{{a, b, c, d}}; }
static inline QM from_cm(CM a, CM b) { return QM{{a.a, a.b, b.a, b.b}}; }
static inline QM operator+(QM x, QM y) {
    return QM{{m_add(x.c[0], y.c[0]), m_add(x.c[1], y.c[1]), m_add(x.c[2], y.c[2]), m_add(x.c[3], y.c[3])}};
}
static inline QM operator-(QM x, QM y) {
    return QM{{m_sub(x.c[0], y.c[0]), m_sub(x.c[1], y.c[1]), m_sub(x.c[2], y.c[2]), m_sub(x.c[3], y.c[3])}};
}
static inline QM operator-(QM x) { return QM{{m_neg(x.c[0]), m_neg(x.c[1]), m_neg(x.c[2]), m_neg(x.c[3])}}; }
static inline QM operator*(QM x, QM y) {
    CM a = x.lo(), b = x.hi(), c = y.lo(), d = y.hi();
    CM bd = b * d;
    CM rbd = {m_sub(m_add(bd.a, bd.a), bd.b), m_add(m_add(bd.b, bd.b), bd.a)};  // (2 + i) * bd
    return from_cm(a * c + rbd, a * d + b * c);
}
static inline QM operator*(QM x, uint32_t m) {
    return QM{{m_mul(x.c[0], m), m_mul(x.c[1], m), m_mul(x.c[2], m), m_mul(x.c[3], m)}};
}
static inline QM mul_cm(QM x, CM o) { return from_cm(x.lo() * o, x.hi() * o); }
static inline QM qm_inv(QM x) {
    CM a = x.lo(), b = x.hi();
    CM b2 = b * b;
    CM ib2 = {m_neg(b2.b), b2.a};
    CM denom = a * a - (b2 + b2 + ib2);
    CM di = cm_inv(denom);
    CM nb = {m_neg(b.a), m_neg(b.b)};
    return from_cm(a * di, nb * di);
}
static inline QM qm_pow(QM b, uint64_t e) {
    QM r = qm(1);
    while (e) {
        if (e & 1) r = r * b;
        b = b * b;
        e >>= 1;
    }
    return r;
}
static inline QM complex_conjugate(QM x) { return QM{{x.c[0], x.c[1], m_neg(x.c[2]), m_neg(x.c[3])}}; }
// SecureField::from_partial_evals: e0 + e1*i + e2*u + e3*iu
static inline QM from_partial_evals(const QM e[4]) {
    return e[0] + e[1] * qm(0, 1, 0, 0) + e[2] * qm(0, 0, 1, 0) + e[3] * qm(0, 0, 0, 1);
}

// ---- packed M31 ------------------------------------------------------------------------------------
#if defined(__AVX512F__) && defined(__AVX512BW__)
#define CPU_W 16
#define CPU_LOGW 4
struct V {
    __m512i v;
};
static inline V vload(const uint32_t* p) { return {_mm512_loadu_si512((const void*)p)}; }
static inline void vstore(uint32_t* p, V a) { _mm512_storeu_si512((void*)p, a.v); }
static inline V vset1(uint32_t x) { return {_mm512_set1_epi32((int)x)}; }
static inline V vzero() { return {_mm512_setzero_si512()}; }
static inline V vadd(V a, V b) {
    __m512i s = _mm512_add_epi32(a.v, b.v);
    return {_mm512_min_epu32(s, _mm512_sub_epi32(s, _mm512_set1_epi32((int)P)))};
}
static inline V vsub(V a, V b) {
    __m512i d = _mm512_sub_epi32(a.v, b.v);
    return {_mm512_min_epu32(d, _mm512_add_epi32(d, _mm512_set1_epi32((int)P)))};
}
static inline V vneg(V a) { return vsub(vzero(), a); }
static inline V vmul(V a, V b) {
    // a * 2b as 64-bit products of the even and the odd lanes: low word >> 1 = (ab mod 2^31), high word = ab >> 31
    __m512i b2 = _mm512_add_epi32(b.v, b.v);
    __m512i pe = _mm512_mul_epu32(a.v, b2);
    __m512i po = _mm512_mul_epu32(_mm512_srli_epi64(a.v, 32), _mm512_srli_epi64(b2, 32));
    __m512i hi = _mm512_mask_blend_epi32(0xAAAA, _mm512_srli_epi64(pe, 32), po);
    __m512i lo = _mm512_srli_epi32(_mm512_mask_blend_epi32(0xAAAA, pe, _mm512_slli_epi64(po, 32)), 1);
    return vadd(V{lo}, V{hi});
}
static inline V vxor(V a, V b) { return {_mm512_xor_si512(a.v, b.v)}; }
static inline V vadd32(V a, V b) { return {_mm512_add_epi32(a.v, b.v)}; }
template <int N>
static inline V vror(V a) {
    return {_mm512_ror_epi32(a.v, N)};
}
// (a || b) -> elements at even positions, elements at odd positions
static inline void vdeinterleave(V a, V b, V& ev, V& od) {
    const __m512i ie = _mm512_setr_epi32(0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30);
    const __m512i io = _mm512_setr_epi32(1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31);
    ev.v = _mm512_permutex2var_epi32(a.v, ie, b.v);
    od.v = _mm512_permutex2var_epi32(a.v, io, b.v);
}
static inline void vinterleave(V ev, V od, V& a, V& b) {
    const __m512i il = _mm512_setr_epi32(0, 16, 1, 17, 2, 18, 3, 19, 4, 20, 5, 21, 6, 22, 7, 23);
    const __m512i ih = _mm512_setr_epi32(8, 24, 9, 25, 10, 26, 11, 27, 12, 28, 13, 29, 14, 30, 15, 31);
    a.v = _mm512_permutex2var_epi32(ev.v, il, od.v);
    b.v = _mm512_permutex2var_epi32(ev.v, ih, od.v);
}
// CNT consecutive values from p, repeated W / CNT times across the vector
template <int CNT>
static inline V vload_pattern(const uint32_t* p) {
    if constexpr (CNT == 16) return vload(p);
    if constexpr (CNT == 8) return {_mm512_broadcast_i64x4(_mm256_loadu_si256((const __m256i*)p))};
    if constexpr (CNT == 4) return {_mm512_broadcast_i32x4(_mm_loadu_si128((const __m128i*)p))};
    if constexpr (CNT == 2) {
        uint64_t x;
        memcpy(&x, p, 8);
        return {_mm512_set1_epi64((long long)x)};
    }
    return vset1(*p);
}
#elif defined(__AVX2__)
#define CPU_W 8
#define CPU_LOGW 3
struct V {
    __m256i v;
};
static inline V vload(const uint32_t* p) { return {_mm256_loadu_si256((const __m256i*)p)}; }
static inline void vstore(uint32_t* p, V a) { _mm256_storeu_si256((__m256i*)p, a.v); }
static inline V vset1(uint32_t x) { return {_mm256_set1_epi32((int)x)}; }
static inline V vzero() { return {_mm256_setzero_si256()}; }
static inline V vadd(V a, V b) {
    __m256i s = _mm256_add_epi32(a.v, b.v);
    return {_mm256_min_epu32(s, _mm256_sub_epi32(s, _mm256_set1_epi32((int)P)))};
}
static inline V vsub(V a, V b) {
    __m256i d = _mm256_sub_epi32(a.v, b.v);
    return {_mm256_min_epu32(d, _mm256_add_epi32(d, _mm256_set1_epi32((int)P)))};
}
static inline V vneg(V a) { return vsub(vzero(), a); }
static inline V vmul(V a, V b) {
    __m256i b2 = _mm256_add_epi32(b.v, b.v);
    __m256i pe = _mm256_mul_epu32(a.v, b2);
    __m256i po = _mm256_mul_epu32(_mm256_srli_epi64(a.v, 32), _mm256_srli_epi64(b2, 32));
    __m256i hi = _mm256_blend_epi32(_mm256_srli_epi64(pe, 32), po, 0xAA);
    __m256i lo = _mm256_srli_epi32(_mm256_blend_epi32(pe, _mm256_slli_epi64(po, 32), 0xAA), 1);
    return vadd(V{lo}, V{hi});
}
static inline V vxor(V a, V b) { return {_mm256_xor_si256(a.v, b.v)}; }
static inline V vadd32(V a, V b) { return {_mm256_add_epi32(a.v, b.v)}; }
template <int N>
static inline V vror(V a) {
    return {_mm256_or_si256(_mm256_srli_epi32(a.v, N), _mm256_slli_epi32(a.v, 32 - N))};
}
static inline void vdeinterleave(V a, V b, V& ev, V& od) {
    __m256 fa = _mm256_castsi256_ps(a.v), fb = _mm256_castsi256_ps(b.v);
    __m256i e = _mm256_castps_si256(_mm256_shuffle_ps(fa, fb, 0x88));  // a0 a2 b0 b2 | a4 a6 b4 b6
    __m256i o = _mm256_castps_si256(_mm256_shuffle_ps(fa, fb, 0xDD));  // a1 a3 b1 b3 | a5 a7 b5 b7
    ev.v = _mm256_permute4x64_epi64(e, 0xD8);
    od.v = _mm256_permute4x64_epi64(o, 0xD8);
}
static inline void vinterleave(V ev, V od, V& a, V& b) {
    __m256i l = _mm256_unpacklo_epi32(ev.v, od.v);  // e0 o0 e1 o1 | e4 o4 e5 o5
    __m256i h = _mm256_unpackhi_epi32(ev.v, od.v);  // e2 o2 e3 o3 | e6 o6 e7 o7
    a.v = _mm256_permute2x128_si256(l, h, 0x20);
    b.v = _mm256_permute2x128_si256(l, h, 0x31);
}
template <int CNT>
static inline V vload_pattern(const uint32_t* p) {
    if constexpr (CNT == 8) return vload(p);
    if constexpr (CNT == 4) return {_mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i*)p))};
    if constexpr (CNT == 2) {
        uint64_t x;
        memcpy(&x, p, 8);
        return {_mm256_set1_epi64x((long long)x)};
    }
    return vset1(*p);
}
#else
#error "cpu_prover needs AVX2 or AVX-512 (build with -march=x86-64-v3 or -march=x86-64-v4)"
#endif
constexpr int W = CPU_W;
constexpr int LOGW = CPU_LOGW;

static inline V operator+(V a, V b) { return vadd(a, b); }
static inline V operator-(V a, V b) { return vsub(a, b); }
static inline V operator*(V a, V b) { return vmul(a, b); }

// packed CM31 / QM31 (structure of vectors)
struct VC {
    V a, b;
};
static inline VC operator+(VC x, VC y) { return {x.a + y.a, x.b + y.b}; }
static inline VC operator-(VC x, VC y) { return {x.a - y.a, x.b - y.b}; }
static inline VC operator*(VC x, VC y) { return {x.a * y.a - x.b * y.b, x.a * y.b + x.b * y.a}; }
static inline VC vc_set1(CM c) { return {vset1(c.a), vset1(c.b)}; }

struct VQ {
    V c[4];
    VC lo() const { return {c[0], c[1]}; }
    VC hi() const { return {c[2], c[3]}; }
};
static inline VQ vq_from(VC a, VC b) { return VQ{{a.a, a.b, b.a, b.b}}; }
static inline VQ vq_zero() { return VQ{{vzero(), vzero(), vzero(), vzero()}}; }
static inline VQ vq_set1(QM q) { return VQ{{vset1(q.c[0]), vset1(q.c[1]), vset1(q.c[2]), vset1(q.c[3])}}; }
static inline VQ vq_from_m31(V x) { return VQ{{x, vzero(), vzero(), vzero()}}; }
static inline VQ operator+(VQ x, VQ y) { return VQ{{x.c[0] + y.c[0], x.c[1] + y.c[1], x.c[2] + y.c[2], x.c[3] + y.c[3]}}; }
static inline VQ operator-(VQ x, VQ y) { return VQ{{x.c[0] - y.c[0], x.c[1] - y.c[1], x.c[2] - y.c[2], x.c[3] - y.c[3]}}; }
static inline VQ operator*(VQ x, VQ y) {
    VC a = x.lo(), b = x.hi(), c = y.lo(), d = y.hi();
    VC bd = b * d;
    VC rbd = {bd.a + bd.a - bd.b, bd.b + bd.b + bd.a};
    return vq_from(a * c + rbd, a * d + b * c);
}
static inline VQ operator*(VQ x, V m) { return VQ{{x.c[0] * m, x.c[1] * m, x.c[2] * m, x.c[3] * m}}; }
static inline VQ vq_mul_cm(VQ x, VC o) { return vq_from(x.lo() * o, x.hi() * o); }
static inline VQ vq_load(const uint32_t* const c[4], size_t i) {
    return VQ{{vload(c[0] + i), vload(c[1] + i), vload(c[2] + i), vload(c[3] + i)}};
}
static inline void vq_store(uint32_t* const c[4], size_t i, VQ q) {
    for (int k = 0; k < 4; k++) vstore(c[k] + i, q.c[k]);
}
static inline QM vq_lane(const VQ& q, int lane) {
    uint32_t t[4][W];
    for (int k = 0; k < 4; k++) vstore(t[k], q.c[k]);
    return QM{{t[0][lane], t[1][lane], t[2][lane], t[3][lane]}};
}

// lane-wise inverses (Fermat; used once per batch-inversion chunk)
static inline V v_inv(V x) {
    // x^(p-2), p-2 = 2^31 - 3: addition chain via squarings
    V r = vset1(1);
    V b = x;
    uint32_t e = P - 2;
    while (e) {
        if (e & 1) r = r * b;
        b = b * b;
        e >>= 1;
    }
    return r;
}
static inline VC vc_inv(VC x) {
    V n = v_inv(x.a * x.a + x.b * x.b);
    return {x.a * n, vneg(x.b) * n};
}
static inline VQ vq_inv(VQ x) {
    VC a = x.lo(), b = x.hi();
    VC b2 = b * b;
    VC ib2 = {vneg(b2.b), b2.a};
    VC denom = a * a - (b2 + b2 + ib2);
    VC di = vc_inv(denom);
    VC nb = {vneg(b.a), vneg(b.b)};
    return vq_from(a * di, nb * di);
}

}  // namespace cpu
