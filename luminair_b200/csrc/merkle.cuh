#pragma once
#include "m31.cuh"

namespace lb {
// out: 2^log_size digests (8 u32 each); prev: 2^(log_size+1) digests or nullptr;
// d_cols: DEVICE array of n_cols device column pointers (each >= 2^log_size u32).
cudaError_t merkle_commit_layer(uint32_t* out, const uint32_t* prev, const uint32_t* const* d_cols, int n_cols,
                                int log_size, cudaStream_t stream);
cudaError_t gather_rows(uint32_t* d_out, const uint32_t* const* d_cols, int n_cols, const uint32_t* d_idx, int n_idx,
                        cudaStream_t stream);
}  // namespace lb
