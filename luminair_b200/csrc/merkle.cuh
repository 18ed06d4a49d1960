#pragma once
#include "m31.cuh"

namespace lb {
// out: 2^log_size digests (8 u32 each); prev: 2^(log_size+1) digests or nullptr;
// d_cols: DEVICE array of n_cols device column pointers (each >= 2^log_size u32).
cudaError_t merkle_commit_layer(uint32_t* out, const uint32_t* prev, const uint32_t* const* d_cols, int n_cols,
                                int log_size, cudaStream_t stream);
constexpr int MERKLE_SMALL_COLS = 16;
struct MerkleColsArg {
    const uint32_t* p[MERKLE_SMALL_COLS];
};
cudaError_t merkle_commit_layer_small(uint32_t* out, const uint32_t* prev, const MerkleColsArg& cols, int n_cols,
                                      int log_size, cudaStream_t stream);
// column-less layers from_log-1 .. 0 in one launch; layers[k] = buffer of layer k (2^k digests), layers[from_log] given
constexpr int MERKLE_TOP_MAX_LOG = 10;
struct MerkleTopArgs {
    uint32_t* layers[MERKLE_TOP_MAX_LOG + 1];
    int from_log;
};
cudaError_t merkle_commit_top(const MerkleTopArgs& args, cudaStream_t stream);
// Several levels per launch for the narrow middle of a tree (2^11 .. 2^17 nodes), where one launch per layer is a chain of
// latency-bound kernels: level 0 = layer `log_top` (children `prev` and / or <= MERKLE_SMALL_COLS columns), then `depth - 1`
// column-less layers above it.  One CTA owns 512 adjacent nodes of layer log_top and what they reduce to.
constexpr int MERKLE_SUBTREE_MAX_DEPTH = 10;
// (threshold re-measured with dependent launch on, cfg-3 proof: 16 -> 7.91 ms, 17 -> 7.81, 18 -> 7.89, 19 -> 8.00)
#ifndef LB_MERKLE_SUBTREE_MAX_LOG
#define LB_MERKLE_SUBTREE_MAX_LOG 17
#endif
constexpr int MERKLE_SUBTREE_MAX_LOG = LB_MERKLE_SUBTREE_MAX_LOG;
struct MerkleSubtreeArgs {
    uint32_t* layers[MERKLE_SUBTREE_MAX_DEPTH];  // layers[d] = buffer of layer log_top - d
    const uint32_t* prev;                        // digests of layer log_top + 1, or nullptr
    MerkleColsArg cols;
    int n_cols, log_top, depth;
};
cudaError_t merkle_commit_subtree(const MerkleSubtreeArgs& args, cudaStream_t stream);
cudaError_t gather_rows(uint32_t* d_out, const uint32_t* const* d_cols, int n_cols, const uint32_t* d_idx, int n_idx,
                        cudaStream_t stream);
}  // namespace lb
