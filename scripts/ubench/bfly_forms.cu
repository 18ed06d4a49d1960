// Register-only throughput of alternative M31 butterfly formulations on sm_100a (B200), next to the shipped one.  Answers
// "which formulation could reach the 0.45 / 0.60 HBM fraction"  with measurements instead of estimates:
// the cfg-2 round trip needs 9.1 / 6.8 clk per warp-butterfly per SM sub-partition ALL-IN; these loops have no memory traffic.
//   F0  shipped: Shoup product (IMAD.WIDE hi + 2 IMAD), umin reductions                       7 instr / butterfly
//   F1  Mersenne fold: one 64-bit product, (lo & P) + (x >> 31) fold on the ALU pipe
//   F2  radix-4 with shared twiddle products: 3 products per 4 outputs (forward direction)
//   F3  FP64: operands as doubles, product by h = a*b, l = fma(a,b,-h), q = rint(h / P), r = fma(-q, P, h) + l
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a bfly_forms.cu -o bfly_forms && ./bfly_forms
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../luminair_b200/csrc/m31.cuh"
using namespace lb;

__device__ __forceinline__ uint32_t rmin(uint32_t x) { return min(x, x - P); }

// ---- F0
__device__ __forceinline__ void f0(uint32_t& v0, uint32_t& v1, uint2 w) {
    uint32_t a = rmin(v0), t = rmin(mul_shoup(v1, w));
    v0 = a + t;
    v1 = a + P - t;
}
// ---- F1: a <= P+1 (reduced), w < P: a * w < 2^62
__device__ __forceinline__ uint32_t mul_fold(uint32_t a, uint32_t w) {
    uint64_t x = (uint64_t)a * w;
    return (uint32_t)(x & P) + (uint32_t)(x >> 31);  // <= 2P
}
__device__ __forceinline__ void f1(uint32_t& v0, uint32_t& v1, uint2 w) {
    uint32_t a = rmin(v0), t = rmin(mul_fold(rmin(v1), w.x));
    v0 = a + t;
    v1 = a + P - t;
}
// ---- F2: forward radix-4 on (a, b, c, d): layer i+1 twiddle w2 for (a,c),(b,d); layer i twiddles w for (a,b), -w for (c,d);
// w3 = w * w2.  a2 = a + P1 + P2 + P3, b2 = a + P1 - P2 - P3, c2 = a - P1 - P2 + P3, d2 = a - P1 + P2 - P3.
__device__ __forceinline__ void f2(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint2 w, uint2 w2, uint2 w3) {
    uint32_t p1 = rmin(mul_shoup(c, w2)), p2 = rmin(mul_shoup(b, w)), p3 = rmin(mul_shoup(d, w3)), ar = rmin(a);
    uint32_t u = rmin(ar + p1), v = rmin(ar + P - p1), s = rmin(p2 + p3), t = rmin(p2 + P - p3);
    a = u + s;
    b = u + P - s;
    c = v + P - t;
    d = v + t;
}
// ---- F3: values as doubles in [0, 2P]
__device__ __forceinline__ double mulmod_d(double a, double w, double invp, double p) {
    double h = a * w, l = fma(a, w, -h);
    double q = rint(h * invp);
    return fma(-q, p, h) + l;  // in (-P, P)
}
__device__ __forceinline__ void f3(double& v0, double& v1, double w, double invp, double p) {
    double t = mulmod_d(v1, w, invp, p);
    double a = v0 - p * floor(v0 * invp);  // reduce the addend
    v0 = a + t;
    v1 = a - t;
}

template <int FORM>
__global__ void __launch_bounds__(256, 3) k(uint32_t* out, const uint2* tw, int iters, long long* cyc) {
    uint32_t v[16];
    uint2 w[8];
    double dv[16], dw[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] = threadIdx.x * 2654435761u + i; dv[i] = (double)(v[i] >> 2); }
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = tw[(threadIdx.x + i) & 63]; dw[i] = (double)w[i].x; }
    const double p = 2147483647.0, invp = 1.0 / 2147483647.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (FORM == 2) {
            // two radix-4 stages = four layers on 16 values: 4 groups of 4 per stage
#pragma unroll
            for (int st = 0; st < 2; ++st)
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int s = st ? 4 : 1;
                    const int base = st ? g : 4 * g;
                    f2(v[base], v[base + s], v[base + 2 * s], v[base + 3 * s], w[g], w[(g + 1) & 7], w[(g + 2) & 7]);
                }
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int pr = 0; pr < 8; ++pr) {
                    const int j0 = ((pr >> b) << (b + 1)) | (pr & ((1 << b) - 1));
                    const int j1 = j0 | (1 << b);
                    if (FORM == 0) f0(v[j0], v[j1], w[pr >> b]);
                    if (FORM == 1) f1(v[j0], v[j1], w[pr >> b]);
                    if (FORM == 3) f3(dv[j0], dv[j1], dw[pr >> b], invp, p);
                }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i] + (uint32_t)(long long)dv[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int FORM>
void run(const char* name, int iters) {
    uint32_t* out; long long* cyc; uint2* tw; long long h;
    const int ctas = 3, blocks = 148 * ctas, threads = 256;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8); cudaMalloc(&tw, 64 * 8);
    uint2 htw[64];
    for (int i = 0; i < 64; ++i) { uint32_t t = 12345u * (i + 1) % P; htw[i] = make_uint2(t, shoup_companion(t)); }
    cudaMemcpy(tw, htw, sizeof(htw), cudaMemcpyHostToDevice);
    k<FORM><<<blocks, threads>>>(out, tw, iters, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<FORM><<<blocks, threads>>>(out, tw, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double wb = (double)iters * 32 * (threads / 32) * ctas / 4.0;  // warp-butterflies per sub-partition (32 butterflies / thread / iter)
    printf("%-46s %.2f clk per warp-butterfly per SMSP by clock64, %.2f by CUDA events at 1.965 GHz (%.3f ms) (%s)\n", name, h / wb,
           ms * 1e-3 * 1.965e9 / wb, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); cudaFree(tw);
}

int main() {
    run<0>("F0 shipped: Shoup + umin", 2000);
    run<1>("F1 Mersenne fold (1 IMAD.WIDE + ALU fold)", 2000);
    run<2>("F2 radix-4, 3 shared products per 4 outputs", 2000);
    run<3>("F3 FP64 (DMUL/DFMA product, floor reduce)", 300);
    printf("needed all-in for the cfg-2 round trip: 9.1 clk (0.45 of HBM), 6.8 clk (0.60)\n");
    return 0;
}
