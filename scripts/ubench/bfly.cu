// Register-only M31 butterfly throughput on sm_100a: the exact butterfly code of cfft.cu on 16 values per thread,
// no memory traffic in the loop.  Tells how far the CFFT kernels are from their arithmetic bound.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../luminair_b200/csrc/m31.cuh"
using namespace lb;

#ifdef RED_MIN
__device__ __forceinline__ uint32_t redm(uint32_t x) { return min(x, x - P); }
#else
__device__ __forceinline__ uint32_t redm(uint32_t x) { return red(x); }
#endif
__device__ __forceinline__ void bf_fwd(uint32_t& v0, uint32_t& v1, uint2 w) {
    uint32_t a = redm(v0);
    uint32_t t = redm(mul_shoup(v1, w));
    v0 = a + t;
    v1 = a + P - t;
}
__device__ __forceinline__ void bf_inv(uint32_t& v0, uint32_t& v1, uint2 w) {
    uint32_t a = redm(v0), b = redm(v1);
    v0 = a + b;
    v1 = mul_shoup(a + P - b, w);
}

template <bool FWD>
__global__ void __launch_bounds__(256, WARPS_CTAS) k(uint32_t* out, const uint2* tw, int iters, long long* cyc) {
    uint32_t v[16];
    uint2 w[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 2654435761u + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = tw[(threadIdx.x + i) & 63];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int pr = 0; pr < 8; ++pr) {
                const int j0 = ((pr >> b) << (b + 1)) | (pr & ((1 << b) - 1));
                const int j1 = j0 | (1 << b);
                if (FWD) bf_fwd(v[j0], v[j1], w[pr >> b]); else bf_inv(v[j0], v[j1], w[pr >> b]);
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    uint32_t* out; long long* cyc; uint2* tw; long long h;
    int blocks = 148 * WARPS_CTAS, threads = 256, iters = 2000;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8); cudaMalloc(&tw, 64 * 8);
    uint2 htw[64];
    for (int i = 0; i < 64; ++i) { uint32_t t = 12345u * (i + 1) % P; htw[i] = make_uint2(t, shoup_companion(t)); }
    cudaMemcpy(tw, htw, sizeof(htw), cudaMemcpyHostToDevice);
    for (int fwd = 0; fwd < 2; ++fwd) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (fwd) k<true><<<blocks, threads>>>(out, tw, iters, cyc); else k<false><<<blocks, threads>>>(out, tw, iters, cyc);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("   kernel %.3f ms -> %.2f clk per warp-butterfly per SMSP at 1.965 GHz\n", ms, ms * 1e-3 * 1.965e9 / ((double)iters * 32 * (threads / 32) * WARPS_CTAS / 4.0));
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double wb = (double)iters * 32 * (threads / 32) * WARPS_CTAS / 4.0;  // warp-butterflies per sub-partition
        printf("%s  CTAs/SM=%d: %lld cycles, %.2f clk per warp-butterfly per SMSP\n", fwd ? "fwd" : "inv", WARPS_CTAS, h, h / wb);
    }
    return 0;
}
