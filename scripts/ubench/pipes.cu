// Integer pipe throughput microbenchmark for sm_100a (B200): warp-instructions per clock per SM
// sub-partition for the instruction classes the M31 butterflies and Blake2s are made of.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipes.cu -o pipes && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NV 8

template <int OP>
__global__ void k(uint32_t* out, uint32_t c0, uint32_t c1, long long* cyc) {
    uint32_t v[NV], w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { v[i] = threadIdx.x * 2654435761u + i + c0; w[i] = threadIdx.x + i * c1; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(c0), "r"(c1));          // IMAD
            if (OP == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(c0));                          // IMAD.HI
            if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(c0), "r"(c1));        // LOP3
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));                           // IADD3 (or IMAD.IADD)
            if (OP == 4) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(v[i]));                             // SHF
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %0, 0x1032;" : "+r"(v[i]));                              // PRMT
            if (OP == 6) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(c0), "r"(c1));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(c0), "r"(c1)); }     // IMAD + LOP3 pair
            if (OP == 7) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(c0));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(c0), "r"(c1)); }     // IMAD.HI + LOP3 pair
            if (OP == 8) { uint32_t t = v[i] >> 31; asm volatile("" :: "r"(t)); v[i] = (v[i] & 0x7FFFFFFFu) + t; v[i] ^= w[i]; }  // red() + xor
            if (OP == 9) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(unsigned long long*)&v[i & ~1]) : "r"(w[i]), "r"(c0)); // IMAD.WIDE
            if (OP == 10) { asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(c0), "r"(c1)); }    // IADD + LOP3 pair
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
    uint32_t* out; long long* cyc; long long h;
    int blocks = 148, threads = 1024;  // 8 warps per sub-partition
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8);
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double winstr = (double)ITERS * NV * per_iter * (threads / 32) / 4.0;  // warp-instr per sub-partition
    printf("%-22s %8lld cycles  %.3f warp-instr/clk/SMSP  (%.2f clk per warp-instr)\n", name, h, winstr / h, h / winstr);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IMAD (mad.lo)", 1);
    run<1>("IMAD.HI (mul.hi)", 1);
    run<2>("LOP3", 1);
    run<3>("IADD (add.u32)", 1);
    run<4>("SHF (rotate)", 1);
    run<5>("PRMT", 1);
    run<6>("IMAD + LOP3", 2);
    run<7>("IMAD.HI + LOP3", 2);
    run<8>("red() + xor (3-4 ALU)", 1);
    run<9>("IMAD.WIDE", 1);
    run<10>("IADD + LOP3", 2);
    return 0;
}
