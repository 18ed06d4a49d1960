// Kernel launches with programmatic dependent launch (PDL, sm_90+): the next kernel of the stream is scheduled while the
// previous one drains instead of after its completion has been signalled and the launch processed.  A proof is ~160 short
// dependent launches (Merkle levels, FRI folds, Fiat-Shamir steps), so the gap between kernels is a visible share of it.
//
// Rules (every kernel of this library follows them):
//   * the FIRST statement of every __global__ function is pdl_wait(): it returns once every kernel the launch depends on has
//     completed and its memory is visible (a no-op for a launch without the attribute).  It comes before any early return, so
//     a grid never completes before its predecessor did (the guarantee is transitive along the stream);
//   * single-CTA, latency-bound kernels (tree tops, Fiat-Shamir steps, the FRI tail, small reductions) call
//     pdl_launch_dependents() right after it: their successor becomes resident at once and sits in its own pdl_wait().
//     Large grids rely on the implicit trigger at block exit (the explicit one in every kernel measured the same: a grid
//     only counts as triggered once ALL its blocks have executed it or exited).
// Measured on B200 (LB_NO_PDL=1 vs default, same build): cfg-3 proof 8.15 -> 7.82 ms, wide 11.12 -> 10.70 ms, cfg-4 MLP
// 3.57 -> 3.18 ms, Add 32x32 1.19 -> 1.05 ms, CFFT round trip of 64 x 2^20 0.576 -> 0.566 ms.
// LB_NO_PDL=1 in the environment launches everything the plain way (A/B measurements, debugging).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace lb {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("LB_NO_PDL");
        return !(e && e[0] && e[0] != '0');
    }();
    return on;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace lb
