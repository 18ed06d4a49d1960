/* CPU restatement of the circle FFT (oracle / cpu_baseline; TEST INFRASTRUCTURE ONLY).
 *
 * Follows stwo prover/backend/cpu/circle.rs @0790eba (un-vendored dependency of LuminAIR,
 * /root/reference/Cargo.toml:21-28): `evaluate` runs layers n-1..1 with x-twiddles then layer 0
 * with y-twiddles using butterfly (v0 + v1 t, v0 - v1 t); `interpolate` runs layers 0..n-1 with
 * ibutterfly (v0 + v1, (v0 - v1) t^-1) and scales by 2^-n.  fft_layer_loop index math:
 * idx0 = (h << (i+1)) + l, idx1 = idx0 + (1 << i).
 *
 * Parity: checked bit-for-bit against oracle/cfft.py (tests/test_oracle_c.py), which is itself
 * pinned by the reference's committed proof.  Threads: OpenMP over columns.
 */
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define P 0x7FFFFFFFu

static inline uint32_t m_mul(uint32_t a, uint32_t b) {
    uint64_t x = (uint64_t)a * b;
    uint32_t r = (uint32_t)(x & P) + (uint32_t)(x >> 31);
    r = (r & P) + (r >> 31);
    return r >= P ? r - P : r;
}
static inline uint32_t m_add(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= P ? s - P : s; }
static inline uint32_t m_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + P - b; }

/* tw[i] = twiddles of layer i (2^(n-1-i) entries) */
static void evaluate_col(uint32_t* v, int n, const uint32_t* const* tw) {
    for (int i = n - 1; i >= 0; --i) {
        size_t half = (size_t)1 << i, nb = (size_t)1 << (n - 1 - i);
        const uint32_t* t = tw[i];
        for (size_t h = 0; h < nb; ++h) {
            uint32_t w = t[h];
            uint32_t* a = v + (h << (i + 1));
            uint32_t* b = a + half;
            for (size_t l = 0; l < half; ++l) {
                uint32_t tmp = m_mul(b[l], w);
                uint32_t x = a[l];
                a[l] = m_add(x, tmp);
                b[l] = m_sub(x, tmp);
            }
        }
    }
}

static void interpolate_col(uint32_t* v, int n, const uint32_t* const* itw, uint32_t inv_n) {
    for (int i = 0; i < n; ++i) {
        size_t half = (size_t)1 << i, nb = (size_t)1 << (n - 1 - i);
        const uint32_t* t = itw[i];
        for (size_t h = 0; h < nb; ++h) {
            uint32_t w = t[h];
            uint32_t* a = v + (h << (i + 1));
            uint32_t* b = a + half;
            for (size_t l = 0; l < half; ++l) {
                uint32_t x = a[l], y = b[l];
                a[l] = m_add(x, y);
                b[l] = m_mul(m_sub(x, y), w);
            }
        }
    }
    size_t N = (size_t)1 << n;
    for (size_t k = 0; k < N; ++k) v[k] = m_mul(v[k], inv_n);
}

/* v: n_cols columns of 2^n u32 (already zero-extended for evaluate), contiguous */
void oracle_cfft_evaluate(uint32_t* v, int n, int n_cols, const uint32_t* const* tw, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < n_cols; ++c) evaluate_col(v + ((size_t)c << n), n, tw);
}

void oracle_cfft_interpolate(uint32_t* v, int n, int n_cols, const uint32_t* const* itw, uint32_t inv_n, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < n_cols; ++c) interpolate_col(v + ((size_t)c << n), n, itw, inv_n);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
