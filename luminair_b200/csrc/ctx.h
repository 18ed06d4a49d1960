// lb_ctx: one per device; owns the stream, the twiddle tables, scratch and the last error text.
#pragma once
#include <string>
#include <vector>

#include "../../include/luminair_b200.h"
#include "blake2s.cuh"
#include "cfft.cuh"

struct lb_comm;

struct lb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // stream-ordered allocations of this context come from its own pool: blocks freed by one context's stream are never handed
    // to another context's stream (the device default pool would make the second stream wait for the first one's free point)
    cudaMemPool_t pool = nullptr;
    bool pool_is_default = false;
    // pinned staging for the small host->device uploads of a call (pointer tables, batch descriptors, channel state): a copy
    // from pageable memory drains the stream before it starts, a copy from here is asynchronous.  One user at a time
    // (host_util.h::Arena), handed on through `ev_stage`.
    uint8_t* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    cudaEvent_t ev_stage = nullptr;
    bool stage_in_flight = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    size_t total_mem = 0;
    lb::Twiddles tw;
    std::string err;
    // scratch for pointer tables / index lists
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    bool kernels_ready = false;
    // host-buffer pipeline (lb_lde_host): copy streams, rotating staging buffers and their events
    static constexpr int PIPE_SLOTS = 3;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[PIPE_SLOTS] = {}, ev_comp[PIPE_SLOTS] = {}, ev_out[PIPE_SLOTS] = {};
    uint32_t* pipe_buf[PIPE_SLOTS] = {};
    size_t pipe_buf_words = 0;
    // diagnostics of the last lb_prove call
    std::vector<lb::Hash32> transcript;  // channel digest after every mix
    std::vector<float> stage_ms;         // wall-clock per stage (stream-synchronised)
};

namespace lb {
// Every extern "C" entry point runs with the context's device current and puts the caller's device back on return: two
// contexts (two GPUs) in one process, or a host that has another device current, must not launch on the wrong device.
struct DeviceGuard {
    int prev = -1;
    bool armed = false;
    explicit DeviceGuard(const lb_ctx* ctx) {
        if (!ctx) return;
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        if (prev != ctx->device) armed = cudaSetDevice(ctx->device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (armed) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
inline int ctx_fail(lb_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        ctx->err = what;
        if (e != cudaSuccess) {
            ctx->err += ": ";
            ctx->err += cudaGetErrorString(e);
        }
    }
    return code;
}
// the prover proper (prover.cu)
int prove_impl(lb_ctx* ctx, const lb_trace_table* tables, int n_tables, const lb_preprocessed_column* pre, int n_pre,
               const lb_prove_config* cfg, std::vector<uint8_t>& out, lb_comm* comm = nullptr);
int eval_at_point_impl(lb_ctx* ctx, const uint32_t* const* h_cols, int n_cols, int log, const uint32_t point[8],
                       uint32_t* h_out);
int accumulate_quotients_impl(lb_ctx* ctx, int log, const uint32_t* const* h_cols, int n_cols,
                              const lb_sample_batch* batches, const lb_batch_shard* shards, int n_batches,
                              const uint32_t random_coeff[4], uint32_t* const d_out[4]);
int fold_impl(lb_ctx* ctx, int circle, uint32_t* const d_dst[4], const uint32_t* const d_src[4], int log,
              const uint32_t alpha[4]);
int grind_impl(lb_ctx* ctx, const uint32_t digest[8], int variant, uint32_t pow_bits, uint64_t* nonce_out);
int logup_impl(lb_ctx* ctx, int kind, const uint32_t* d_main, size_t main_stride, const uint32_t* const d_lut[2],
               uint32_t* d_inter, size_t inter_stride, int log, const lb_relation* rels, int n_rels, uint32_t claimed_out[4]);
int constraint_quotients_impl(lb_ctx* ctx, int kind, const uint32_t* d_main, size_t main_stride, const uint32_t* d_inter,
                              size_t inter_stride, const uint32_t* const d_lut[2], int log_size, int eval_log,
                              const lb_relation* rels, int n_rels, const uint32_t claimed_sum[4], const uint32_t* pows,
                              int n_pows, uint32_t* const d_acc[4], int accumulate);
}  // namespace lb
