// M31 / CM31 / QM31 device arithmetic for sm_100a.
//
// Field definitions follow stwo core/fields/{m31,cm31,qm31}.rs (un-vendored dependency of
// /root/reference, Cargo.toml:21-28).  The instruction selection is B200-specific:
//   * twiddle multiplications use Shoup's precomputed-quotient form: 3 IMAD-class
//     instructions on the FMA pipe, zero ALU-pipe work, any u32 multiplicand, result in [0,2P)
//   * sums stay lazy u32 representatives, folded with (x & P) + (x >> 31) only where the
//     next add could overflow (2P+1 = 2^32-1: Mersenne-31 has no spare bit)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lb {

constexpr uint32_t P = 0x7FFFFFFFu;

// ---- lazy helpers -----------------------------------------------------------------
// any u32 -> [0, P+1]  (<= P when x <= 2P)
__device__ __forceinline__ uint32_t red(uint32_t x) { return (x & P) + (x >> 31); }
// [0, 2P) -> [0, P)
__device__ __forceinline__ uint32_t canon2(uint32_t x) { return min(x, x - P); }
// any u32 -> [0, P)
__device__ __forceinline__ uint32_t canon(uint32_t x) {
    x = red(x);
    return min(x, x - P);
}

// Shoup multiplication: w = (t, floor(t * 2^32 / P)), t < P, a any u32 -> a*t mod P in [0, 2P)
// The quotient word comes from a 64-bit IMAD.WIDE rather than IMAD.HI: measured on B200
// (scripts/ubench/pipes.cu) IMAD.WIDE issues every 2.3 clk per SM sub-partition, IMAD.HI every 4.
#ifndef LB_MULHI_WIDE
#define LB_MULHI_WIDE 1
#endif
__device__ __forceinline__ uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if LB_MULHI_WIDE
    uint32_t lo, hi;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
    (void)lo;
    return hi;
#else
    return __umulhi(a, b);
#endif
}
__device__ __forceinline__ uint32_t mul_shoup(uint32_t a, uint2 w) {
    uint32_t q = mulhi_u32(a, w.y);
    return a * w.x - q * P;
}
__host__ __device__ __forceinline__ uint32_t shoup_companion(uint32_t t) {
    return (uint32_t)(((uint64_t)t << 32) / P);
}

// ---- canonical M31 ([0,P) in, [0,P) out) -------------------------------------------
__host__ __device__ __forceinline__ uint32_t m_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return s >= P ? s - P : s;
}
__host__ __device__ __forceinline__ uint32_t m_sub(uint32_t a, uint32_t b) {
    return a >= b ? a - b : a + P - b;
}
__host__ __device__ __forceinline__ uint32_t m_neg(uint32_t a) { return a ? P - a : 0; }
__host__ __device__ __forceinline__ uint32_t m_reduce64(uint64_t x) {
    // x < 2^62
    uint32_t r = (uint32_t)(x & P) + (uint32_t)(x >> 31);  // < 2^32
    r = (r & P) + (r >> 31);
    return r >= P ? r - P : r;
}
__host__ __device__ __forceinline__ uint32_t m_mul(uint32_t a, uint32_t b) {
    return m_reduce64((uint64_t)a * b);
}
__host__ __device__ inline uint32_t m_pow(uint32_t a, uint32_t e) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = m_mul(r, a);
        a = m_mul(a, a);
        e >>= 1;
    }
    return r;
}
// a^(p-2) with the addition chain for 2^31 - 3 (30 squarings + 7 products instead of 30 + 29 for square-and-multiply)
__host__ __device__ __forceinline__ uint32_t m_sqn(uint32_t a, int n) {
    for (int i = 0; i < n; ++i) a = m_mul(a, a);
    return a;
}
__host__ __device__ inline uint32_t m_inv(uint32_t a) {
    uint32_t t0 = m_mul(m_sqn(a, 2), a);     // a^5
    uint32_t t1 = m_mul(m_sqn(t0, 1), t0);   // a^15
    uint32_t t2 = m_mul(m_sqn(t1, 3), t0);   // a^125
    uint32_t t3 = m_mul(m_sqn(t2, 1), t0);   // a^255
    uint32_t t4 = m_mul(m_sqn(t3, 8), t3);   // a^(2^16 - 1)
    uint32_t t5 = m_mul(m_sqn(t4, 8), t3);   // a^(2^24 - 1)
    return m_mul(m_sqn(t5, 7), t2);          // a^(2^31 - 3)
}

// ---- CM31 -------------------------------------------------------------------------
struct CM31 {
    uint32_t a, b;
};
__host__ __device__ __forceinline__ CM31 c_add(CM31 x, CM31 y) { return {m_add(x.a, y.a), m_add(x.b, y.b)}; }
__host__ __device__ __forceinline__ CM31 c_sub(CM31 x, CM31 y) { return {m_sub(x.a, y.a), m_sub(x.b, y.b)}; }
__host__ __device__ __forceinline__ CM31 c_neg(CM31 x) { return {m_neg(x.a), m_neg(x.b)}; }
// x < 2^63 -> canonical M31
__host__ __device__ __forceinline__ uint32_t m_reduce63(uint64_t x) {
    uint64_t y = (x & P) + (x >> 31);                       // < 3 * 2^31
    uint32_t z = (uint32_t)(y & P) + (uint32_t)(y >> 31);   // <= P + 2
    return z >= P ? z - P : z;
}
__host__ __device__ __forceinline__ CM31 c_mul(CM31 x, CM31 y) {
    // (a+bi)(c+di) = (ac - bd) + (ad + bc) i.  Each product is < 2^62, so both coordinates are accumulated as one
    // 64-bit sum (-bd as (P - b) d) and reduced once: two reductions per product instead of four.
    uint64_t re = (uint64_t)x.a * y.a + (uint64_t)(P - x.b) * y.b;
    uint64_t im = (uint64_t)x.a * y.b + (uint64_t)x.b * y.a;
    return {m_reduce63(re), m_reduce63(im)};
}
__host__ __device__ __forceinline__ CM31 c_mul_m(CM31 x, uint32_t s) { return {m_mul(x.a, s), m_mul(x.b, s)}; }
__host__ __device__ inline CM31 c_inv(CM31 x) {
    uint32_t n = m_inv(m_add(m_mul(x.a, x.a), m_mul(x.b, x.b)));
    return {m_mul(x.a, n), m_mul(m_neg(x.b), n)};
}

// ---- QM31 = CM31[u]/(u^2 - (2+i)) ---------------------------------------------------
struct QM31 {
    CM31 a, b;
};
__host__ __device__ __forceinline__ QM31 q_make(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return {{c0, c1}, {c2, c3}};
}
__host__ __device__ __forceinline__ QM31 q_zero() { return {{0, 0}, {0, 0}}; }
__host__ __device__ __forceinline__ QM31 q_from_m(uint32_t x) { return {{x, 0}, {0, 0}}; }
__host__ __device__ __forceinline__ QM31 q_add(QM31 x, QM31 y) { return {c_add(x.a, y.a), c_add(x.b, y.b)}; }
__host__ __device__ __forceinline__ QM31 q_sub(QM31 x, QM31 y) { return {c_sub(x.a, y.a), c_sub(x.b, y.b)}; }
__host__ __device__ __forceinline__ QM31 q_neg(QM31 x) { return {c_neg(x.a), c_neg(x.b)}; }
__host__ __device__ __forceinline__ CM31 c_mul_r(CM31 x) {
    // (2 + i) * (a + bi) = (2a - b) + (a + 2b) i
    return {m_sub(m_add(x.a, x.a), x.b), m_add(m_add(x.b, x.b), x.a)};
}
__host__ __device__ __forceinline__ QM31 q_mul(QM31 x, QM31 y) {
    CM31 ac = c_mul(x.a, y.a), bd = c_mul(x.b, y.b);
    CM31 ad = c_mul(x.a, y.b), bc = c_mul(x.b, y.a);
    return {c_add(ac, c_mul_r(bd)), c_add(ad, bc)};
}
__host__ __device__ __forceinline__ QM31 q_mul_m(QM31 x, uint32_t s) { return {c_mul_m(x.a, s), c_mul_m(x.b, s)}; }
__host__ __device__ __forceinline__ QM31 q_mul_c(QM31 x, CM31 s) { return {c_mul(x.a, s), c_mul(x.b, s)}; }
__host__ __device__ inline QM31 q_inv(QM31 x) {
    CM31 b2 = c_mul(x.b, x.b);
    CM31 denom = c_sub(c_mul(x.a, x.a), c_mul_r(b2));
    CM31 di = c_inv(denom);
    return {c_mul(x.a, di), c_mul(c_neg(x.b), di)};
}
__host__ __device__ inline QM31 q_pow(QM31 x, uint64_t e) {
    QM31 r = q_from_m(1);
    while (e) {
        if (e & 1) r = q_mul(r, x);
        x = q_mul(x, x);
        e >>= 1;
    }
    return r;
}
__host__ __device__ __forceinline__ bool q_eq(QM31 x, QM31 y) {
    return x.a.a == y.a.a && x.a.b == y.a.b && x.b.a == y.b.a && x.b.b == y.b.b;
}

// ---- circle group --------------------------------------------------------------------
struct Pt {
    uint32_t x, y;
};
__host__ __device__ __forceinline__ Pt pt_add(Pt p, Pt q) {
    return {m_sub(m_mul(p.x, q.x), m_mul(p.y, q.y)), m_add(m_mul(p.x, q.y), m_mul(p.y, q.x))};
}

}  // namespace lb
