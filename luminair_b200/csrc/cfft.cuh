#pragma once
#include "m31.cuh"

namespace lb {

// Twiddle tables shared by every domain of log size <= max_log (see cfft.cu).
// Entries are Shoup pairs (t, floor(t*2^32/P)).
struct Twiddles {
    int max_log = 0;
    uint2* fwd = nullptr;  // X[1..K] then Y[1..K], K = max_log-1
    uint2* inv = nullptr;  // same layout, inverses
    size_t y_off = 0;
};

cudaError_t twiddles_create(Twiddles* tw, int max_log, cudaStream_t stream);
void twiddles_destroy(Twiddles* tw);
cudaError_t twiddles_export_stwo(const Twiddles* tw, int root_log, uint32_t* d_out, cudaStream_t stream);

// values (bit-reversed circle-domain order) -> coefficients, in place, n_cols columns at `stride`
cudaError_t cfft_interpolate(const Twiddles* tw, uint32_t* data, size_t stride, int n_cols, int log_n, int sm_count,
                             cudaStream_t stream);
// coefficients (2^log_in) -> values on CanonicCoset(log_out).circle_domain()
cudaError_t cfft_evaluate(const Twiddles* tw, const uint32_t* coeffs, size_t src_stride, int log_in, uint32_t* out,
                          size_t dst_stride, int log_out, int n_cols, int sm_count, cudaStream_t stream);

// same, but the final pass stores row shard s (rows [s*R/W, (s+1)*R/W), R = 2^log_out, W = n_peers) of column c into
// peers[s] + (peer_col0 + c) * (R/W): the column->row all-to-all of a sharded commit fused into the transform.
// `out` is still needed as the scratch of the earlier passes.  Requires log_out >= 16 and R/W >= 4096.
cudaError_t cfft_evaluate_scatter(const Twiddles* tw, const uint32_t* coeffs, size_t src_stride, int log_in, uint32_t* out,
                                  size_t dst_stride, int log_out, int n_cols, uint32_t* const* peers, int n_peers,
                                  size_t peer_col0, int sm_count, cudaStream_t stream);

}  // namespace lb
