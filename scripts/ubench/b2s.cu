// Register-only Blake2s compression throughput on sm_100a: how far the Merkle kernels are from the pipe bound,
// and what instruction mix / ILP gets closest.  One or two independent compressions per thread, message in registers.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../luminair_b200/csrc b2s.cu -o b2s && ./b2s
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../luminair_b200/csrc/blake2s.cuh"
using namespace lb;

template <int MODE, int NH>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t one, int iters) {
    uint32_t h[NH][8], m[NH][16];
#pragma unroll
    for (int q = 0; q < NH; ++q) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h[q][i] = threadIdx.x * 2654435761u + i + q;
#pragma unroll
        for (int i = 0; i < 16; ++i) m[q][i] = blockIdx.x * 40503u + i * 77u + q;
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < NH; ++q) {
            if (MODE == 0) blake2s_compress(h[q], m[q], 64, 0, 0);
            else blake2s_compress_fma(h[q], m[q], 64, 0, 0, one);
            m[q][it & 15] ^= h[q][0];
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < NH; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) s += h[q][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NH>
void run(const char* name, int ctas_per_sm) {
    uint32_t* out;
    int blocks = 148 * ctas_per_sm, threads = 256, iters = 400;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE, NH><<<blocks, threads>>>(out, 1, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
    }
    double n = (double)blocks * threads * iters * NH;
    printf("%-34s CTAs/SM %d: %.3f ms  %.2f G compressions/s\n", name, ctas_per_sm, ms, n / ms * 1e-6);
    cudaFree(out);
}

int main() {
    for (int c : {2, 4, 6, 8}) {
        run<0, 1>("IADD3 adds, 1 hash/thread", c);
        run<1, 1>("IMAD adds, 1 hash/thread", c);
    }
    for (int c : {2, 4}) {
        run<0, 2>("IADD3 adds, 2 hashes/thread", c);
        run<1, 2>("IMAD adds, 2 hashes/thread", c);
    }
    return 0;
}
