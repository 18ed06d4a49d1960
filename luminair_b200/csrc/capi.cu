// extern "C" boundary of libluminair_b200.so — see include/luminair_b200.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "comm.h"
#include "ctx.h"
#include "kernels.cuh"
#include "merkle.cuh"

namespace {
int fail(lb_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) { return lb::ctx_fail(ctx, code, what, e); }
#define CK(call, what)                                                          \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess)                                                  \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? LB_ERR_OOM : LB_ERR_CUDA, what, _e); \
    } while (0)

int ensure_scratch(lb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return LB_OK;
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes < 65536 ? 65536 : bytes;
    CK(cudaMalloc(&ctx->d_scratch, want), "scratch alloc");
    ctx->scratch_bytes = want;
    return LB_OK;
}
}  // namespace

extern "C" {

int lb_version(void) { return 2; }

int lb_ctx_create(int device, lb_ctx** out) {
    if (!out) return LB_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return LB_ERR_CUDA;
    lb_ctx* ctx = new lb_ctx();
    ctx->device = device;
    int caller_device = -1;
    cudaGetDevice(&caller_device);
    struct Restore {
        int d;
        ~Restore() {
            if (d >= 0) cudaSetDevice(d);
        }
    } restore{caller_device};  // the calling thread keeps its current device
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete ctx;
        return LB_ERR_CUDA;
    }
    {
        // a pool per context, its memory cached between prove() calls
        cudaMemPoolProps pp = {};
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        if (cudaMemPoolCreate(&ctx->pool, &pp) != cudaSuccess) {
            cudaGetLastError();
            ctx->pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&ctx->pool, device) != cudaSuccess) {
                cudaStreamDestroy(ctx->stream);
                delete ctx;
                return LB_ERR_CUDA;
            }
            ctx->pool_is_default = true;
        }
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    {
        void* hp = nullptr;
        const size_t bytes = (size_t)2 << 20;
        if (cudaHostAlloc(&hp, bytes, cudaHostAllocDefault) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming) == cudaSuccess) {
            ctx->h_stage = (uint8_t*)hp;
            ctx->h_stage_bytes = bytes;
        } else {  // uploads fall back to pageable copies
            cudaGetLastError();
            if (hp) cudaFreeHost(hp);
        }
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = prop.totalGlobalMem;
    *out = ctx;
    return LB_OK;
}

void lb_ctx_destroy(lb_ctx* ctx) {
    if (!ctx) return;
    lb::DeviceGuard _dg(ctx);
    cudaStreamSynchronize(ctx->stream);
    lb::twiddles_destroy(&ctx->tw);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    for (int i = 0; i < lb_ctx::PIPE_SLOTS; ++i) {
        if (ctx->pipe_buf[i]) cudaFree(ctx->pipe_buf[i]);
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaStreamDestroy(ctx->stream);
    // blocks still held by the caller (lb_alloc_pooled) keep the pool alive until they are freed
    if (ctx->pool && !ctx->pool_is_default) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

const char* lb_last_error(lb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int lb_sync(lb_ctx* ctx) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    return LB_OK;
}

int lb_alloc(lb_ctx* ctx, size_t n_u32, uint32_t** d_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_out) return LB_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    void* p = nullptr;
    CK(cudaMalloc(&p, (n_u32 ? n_u32 : 1) * sizeof(uint32_t)), "alloc");
    *d_out = (uint32_t*)p;
    return LB_OK;
}

int lb_free(lb_ctx* ctx, uint32_t* d_ptr) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream), "free/sync");
    CK(cudaFree(d_ptr), "free");
    return LB_OK;
}

// stream-ordered variants (cudaMallocFromPoolAsync on the context's stream, memory kept in the context's pool): no device-wide
// synchronisation, for short-lived buffers such as the tensors and tables of a device-side gen_trace.  Not for lb_ipc_export.
int lb_alloc_pooled(lb_ctx* ctx, size_t n_u32, uint32_t** d_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_out) return LB_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    void* p = nullptr;
    CK(cudaMallocFromPoolAsync(&p, (n_u32 ? n_u32 : 1) * sizeof(uint32_t), ctx->pool, ctx->stream), "alloc_pooled");
    *d_out = (uint32_t*)p;
    return LB_OK;
}

int lb_free_pooled(lb_ctx* ctx, uint32_t* d_ptr) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaFreeAsync(d_ptr, ctx->stream), "free_pooled");
    return LB_OK;
}

int lb_memset_zero(lb_ctx* ctx, uint32_t* d_ptr, size_t n_u32) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaMemsetAsync(d_ptr, 0, n_u32 * sizeof(uint32_t), ctx->stream), "memset");
    return LB_OK;
}

int lb_upload(lb_ctx* ctx, uint32_t* d_dst, const uint32_t* h_src, size_t n_u32) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || (!h_src && n_u32)) return LB_ERR_BAD_ARG;
    CK(cudaMemcpyAsync(d_dst, h_src, n_u32 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream), "upload");
    CK(cudaStreamSynchronize(ctx->stream), "upload/sync");
    return LB_OK;
}

int lb_download(lb_ctx* ctx, uint32_t* h_dst, const uint32_t* d_src, size_t n_u32) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || (!h_dst && n_u32)) return LB_ERR_BAD_ARG;
    CK(cudaMemcpyAsync(h_dst, d_src, n_u32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream), "download");
    CK(cudaStreamSynchronize(ctx->stream), "download/sync");
    return LB_OK;
}

int lb_copy(lb_ctx* ctx, uint32_t* d_dst, const uint32_t* d_src, size_t n_u32) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaMemcpyAsync(d_dst, d_src, n_u32 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream), "copy");
    return LB_OK;
}

int lb_timer_start(lb_ctx* ctx) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaEventRecord(ctx->ev0, ctx->stream), "timer start");
    return LB_OK;
}

int lb_timer_stop_ms(lb_ctx* ctx, float* ms_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !ms_out) return LB_ERR_BAD_ARG;
    CK(cudaEventRecord(ctx->ev1, ctx->stream), "timer stop");
    CK(cudaEventSynchronize(ctx->ev1), "timer sync");
    CK(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1), "timer elapsed");
    return LB_OK;
}

int lb_device_info(lb_ctx* ctx, int* sm_count, size_t* total_mem_bytes) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (total_mem_bytes) *total_mem_bytes = ctx->total_mem;
    return LB_OK;
}

int lb_twiddles_ensure(lb_ctx* ctx, int max_log) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || max_log < 1 || max_log > 28) return fail(ctx, LB_ERR_BAD_ARG, "twiddles: max_log out of range");
    if (max_log < 2) max_log = 2;
    if (ctx->tw.max_log >= max_log) return LB_OK;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream), "twiddles/sync");
    lb::twiddles_destroy(&ctx->tw);
    ctx->tw.max_log = 0;
    CK(lb::twiddles_create(&ctx->tw, max_log, ctx->stream), "twiddles create");
    return LB_OK;
}

int lb_twiddles_export(lb_ctx* ctx, int root_log, uint32_t* d_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_out) return LB_ERR_BAD_ARG;
    if (ctx->tw.max_log < root_log + 1) return fail(ctx, LB_ERR_BAD_ARG, "twiddles_export: tables too small");
    CK(lb::twiddles_export_stwo(&ctx->tw, root_log, d_out, ctx->stream), "twiddles export");
    return LB_OK;
}

int lb_interpolate_batch(lb_ctx* ctx, uint32_t* d_cols, size_t stride, int n_cols, int log_size) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || n_cols < 0 || log_size < 1) return fail(ctx, LB_ERR_BAD_ARG, "interpolate: bad args");
    if (stride < ((size_t)1 << log_size) && n_cols > 1) return fail(ctx, LB_ERR_BAD_ARG, "interpolate: stride < column size");
    int r = lb_twiddles_ensure(ctx, log_size);
    if (r) return r;
    CK(lb::cfft_interpolate(&ctx->tw, d_cols, stride, n_cols, log_size, ctx->sm_count, ctx->stream), "interpolate");
    return LB_OK;
}

int lb_evaluate_batch(lb_ctx* ctx, const uint32_t* d_coeffs, size_t src_stride, int log_in, uint32_t* d_out,
                      size_t dst_stride, int log_out, int n_cols) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || n_cols < 0 || log_out < 1 || log_in < 0 || log_in > log_out)
        return fail(ctx, LB_ERR_BAD_ARG, "evaluate: bad args");
    int r = lb_twiddles_ensure(ctx, log_out);
    if (r) return r;
    CK(lb::cfft_evaluate(&ctx->tw, d_coeffs, src_stride, log_in, d_out, dst_stride, log_out, n_cols, ctx->sm_count,
                         ctx->stream),
       "evaluate");
    return LB_OK;
}

int lb_merkle_commit_layer(lb_ctx* ctx, int log_size, const uint32_t* d_prev, const uint32_t* const* h_cols,
                           int n_cols, uint32_t* d_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_out || log_size < 0 || log_size > 30 || n_cols < 0 || (n_cols && !h_cols))
        return fail(ctx, LB_ERR_BAD_ARG, "merkle: bad args");
    const uint32_t* const* d_cols = nullptr;
    if (n_cols) {
        int r = ensure_scratch(ctx, n_cols * sizeof(void*));
        if (r) return r;
        // the table must outlive earlier kernels still reading the scratch
        CK(cudaStreamSynchronize(ctx->stream), "merkle/sync");
        CK(cudaMemcpyAsync(ctx->d_scratch, h_cols, n_cols * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream),
           "merkle/ptrs");
        d_cols = (const uint32_t* const*)ctx->d_scratch;
    }
    CK(lb::merkle_commit_layer(d_out, d_prev, d_cols, n_cols, log_size, ctx->stream), "merkle layer");
    return LB_OK;
}

int lb_gather_rows(lb_ctx* ctx, const uint32_t* const* h_cols, int n_cols, const uint32_t* h_idx, int n_idx,
                   uint32_t* h_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || n_cols < 0 || n_idx < 0) return fail(ctx, LB_ERR_BAD_ARG, "gather: bad args");
    if (n_cols == 0 || n_idx == 0) return LB_OK;
    if (!h_cols || !h_idx || !h_out) return fail(ctx, LB_ERR_BAD_ARG, "gather: null pointer");
    size_t pb = (size_t)n_cols * sizeof(void*), ib = (size_t)n_idx * 4, ob = (size_t)n_cols * n_idx * 4;
    size_t pb_al = (pb + 255) & ~(size_t)255, ib_al = (ib + 255) & ~(size_t)255;
    int r = ensure_scratch(ctx, pb_al + ib_al + ob);
    if (r) return r;
    CK(cudaStreamSynchronize(ctx->stream), "gather/sync");
    char* base = (char*)ctx->d_scratch;
    CK(cudaMemcpyAsync(base, h_cols, pb, cudaMemcpyHostToDevice, ctx->stream), "gather/ptrs");
    CK(cudaMemcpyAsync(base + pb_al, h_idx, ib, cudaMemcpyHostToDevice, ctx->stream), "gather/idx");
    uint32_t* d_out = (uint32_t*)(base + pb_al + ib_al);
    CK(lb::gather_rows(d_out, (const uint32_t* const*)base, n_cols, (const uint32_t*)(base + pb_al), n_idx, ctx->stream),
       "gather");
    CK(cudaMemcpyAsync(h_out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream), "gather/out");
    CK(cudaStreamSynchronize(ctx->stream), "gather/sync2");
    return LB_OK;
}

int lb_eval_at_point(lb_ctx* ctx, const uint32_t* const* h_cols, int n_cols, int log_size, const uint32_t point[8],
                     uint32_t* h_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || n_cols < 0 || log_size < 0 || log_size > 30 || (n_cols && (!h_cols || !h_out)) || !point)
        return fail(ctx, LB_ERR_BAD_ARG, "eval_at_point: bad args");
    if (n_cols == 0) return LB_OK;
    return lb::eval_at_point_impl(ctx, h_cols, n_cols, log_size, point, h_out);
}

int lb_accumulate_quotients(lb_ctx* ctx, int log_size, const uint32_t* const* h_cols, int n_cols,
                            const lb_sample_batch* batches, int n_batches, const uint32_t random_coeff[4],
                            uint32_t* const d_out[4]) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !h_cols || !batches || !random_coeff || !d_out || n_cols < 1 || n_batches < 1 || log_size < 1 || log_size > 30)
        return fail(ctx, LB_ERR_BAD_ARG, "accumulate_quotients: bad args");
    return lb::accumulate_quotients_impl(ctx, log_size, h_cols, n_cols, batches, nullptr, n_batches, random_coeff, d_out);
}

int lb_accumulate_quotients_shard(lb_ctx* ctx, int log_size, const uint32_t* const* h_cols, int n_cols,
                                  const lb_sample_batch* batches, const lb_batch_shard* shards, int n_batches,
                                  const uint32_t random_coeff[4], uint32_t* const d_out[4]) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !h_cols || !batches || !shards || !random_coeff || !d_out || n_cols < 1 || n_batches < 1 || log_size < 1 ||
        log_size > 30)
        return fail(ctx, LB_ERR_BAD_ARG, "accumulate_quotients_shard: bad args");
    return lb::accumulate_quotients_impl(ctx, log_size, h_cols, n_cols, batches, shards, n_batches, random_coeff, d_out);
}

int lb_fold_circle_into_line(lb_ctx* ctx, uint32_t* const d_dst[4], const uint32_t* const d_src[4], int log_size,
                             const uint32_t alpha[4]) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_dst || !d_src || !alpha) return fail(ctx, LB_ERR_BAD_ARG, "fold_circle_into_line: bad args");
    return lb::fold_impl(ctx, 1, d_dst, d_src, log_size, alpha);
}

int lb_fold_line(lb_ctx* ctx, uint32_t* const d_dst[4], const uint32_t* const d_src[4], int log_size, const uint32_t alpha[4]) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_dst || !d_src || !alpha) return fail(ctx, LB_ERR_BAD_ARG, "fold_line: bad args");
    return lb::fold_impl(ctx, 0, d_dst, d_src, log_size, alpha);
}

int lb_grind(lb_ctx* ctx, const uint32_t digest[8], int channel_variant, uint32_t pow_bits, uint64_t* nonce_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !digest || !nonce_out || channel_variant < 0 || channel_variant > 1)
        return fail(ctx, LB_ERR_BAD_ARG, "grind: bad args");
    return lb::grind_impl(ctx, digest, channel_variant, pow_bits, nonce_out);
}

int lb_logup_interaction_trace(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride, uint32_t* d_inter,
                               size_t inter_stride, int log_size, const uint32_t z[4], const uint32_t alpha[4],
                               uint32_t claimed_out[4]) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_main || !d_inter || !z || !alpha || !claimed_out || log_size < 1 || log_size > 30)
        return fail(ctx, LB_ERR_BAD_ARG, "logup: bad args");
    lb_relation node;
    std::memcpy(node.z, z, 16);
    std::memcpy(node.alpha, alpha, 16);
    return lb::logup_impl(ctx, component, d_main, main_stride, nullptr, d_inter, inter_stride, log_size, &node, 1, claimed_out);
}

int lb_logup_interaction_trace_lut(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride,
                                   const uint32_t* const d_lut[2], uint32_t* d_inter, size_t inter_stride, int log_size,
                                   const lb_relation rels[LB_REL_COUNT], uint32_t claimed_out[4]) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_main || !d_inter || !rels || !claimed_out || log_size < 1 || log_size > 30)
        return fail(ctx, LB_ERR_BAD_ARG, "logup: bad args");
    return lb::logup_impl(ctx, component, d_main, main_stride, d_lut, d_inter, inter_stride, log_size, rels, LB_REL_COUNT,
                          claimed_out);
}

int lb_constraint_quotients(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride, const uint32_t* d_inter,
                            size_t inter_stride, int log_size, const uint32_t z[4], const uint32_t alpha[4],
                            const uint32_t claimed_sum[4], const uint32_t* pows, int n_pows, uint32_t* const d_acc[4],
                            int accumulate) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_main || !d_inter || !z || !alpha || !claimed_sum || !pows || !d_acc || log_size < 1 || log_size > 29)
        return fail(ctx, LB_ERR_BAD_ARG, "constraint_quotients: bad args");
    lb_relation node;
    std::memcpy(node.z, z, 16);
    std::memcpy(node.alpha, alpha, 16);
    return lb::constraint_quotients_impl(ctx, component, d_main, main_stride, d_inter, inter_stride, nullptr, log_size,
                                         log_size + 1, &node, 1, claimed_sum, pows, n_pows, d_acc, accumulate);
}

int lb_constraint_quotients_lut(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride, const uint32_t* d_inter,
                                size_t inter_stride, const uint32_t* const d_lut[2], int log_size, int eval_log_size,
                                const lb_relation rels[LB_REL_COUNT], const uint32_t claimed_sum[4], const uint32_t* pows,
                                int n_pows, uint32_t* const d_acc[4], int accumulate) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_main || !d_inter || !rels || !claimed_sum || !pows || !d_acc || log_size < 1 || log_size > 29)
        return fail(ctx, LB_ERR_BAD_ARG, "constraint_quotients: bad args");
    return lb::constraint_quotients_impl(ctx, component, d_main, main_stride, d_inter, inter_stride, d_lut, log_size,
                                         eval_log_size, rels, LB_REL_COUNT, claimed_sum, pows, n_pows, d_acc, accumulate);
}

int lb_evaluate_batch_scatter(lb_ctx* ctx, const uint32_t* d_coeffs, size_t src_stride, int log_in, uint32_t* d_scratch,
                              size_t dst_stride, int log_out, int n_cols, uint32_t* const* h_peers, int n_peers,
                              size_t peer_col0) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || n_cols < 0 || log_out < 1 || log_in < 0 || log_in > log_out || n_peers < 1 || n_peers > 8 || !h_peers)
        return fail(ctx, LB_ERR_BAD_ARG, "evaluate_scatter: bad args");
    int r = lb_twiddles_ensure(ctx, log_out);
    if (r) return r;
    CK(lb::cfft_evaluate_scatter(&ctx->tw, d_coeffs, src_stride, log_in, d_scratch, dst_stride, log_out, n_cols, h_peers, n_peers,
                                 peer_col0, ctx->sm_count, ctx->stream),
       "evaluate_scatter");
    return LB_OK;
}

int lb_ipc_export(lb_ctx* ctx, const uint32_t* d_ptr, uint8_t handle_out[64]) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_ptr || !handle_out) return fail(ctx, LB_ERR_BAD_ARG, "ipc_export: bad args");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaSetDevice(ctx->device);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, (void*)d_ptr), "ipc export");
    std::memcpy(handle_out, &h, 64);
    return LB_OK;
}

int lb_ipc_open(lb_ctx* ctx, const uint8_t handle[64], uint32_t** d_ptr_out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !handle || !d_ptr_out) return fail(ctx, LB_ERR_BAD_ARG, "ipc_open: bad args");
    cudaSetDevice(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "ipc open");
    *d_ptr_out = (uint32_t*)p;
    return LB_OK;
}

int lb_ipc_close(lb_ctx* ctx, uint32_t* d_ptr) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx) return LB_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream), "ipc close/sync");
    CK(cudaIpcCloseMemHandle(d_ptr), "ipc close");
    return LB_OK;
}

int lb_lde_host(lb_ctx* ctx, const uint32_t* h_values, uint32_t* h_evals, int n_cols, int log_in, int log_out,
                uint32_t* h_coeffs, int chunk_cols) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !h_values || !h_evals || n_cols < 0 || log_in < 1 || log_out < log_in || log_out > 28)
        return fail(ctx, LB_ERR_BAD_ARG, "lde_host: bad args");
    if (n_cols == 0) return LB_OK;
    cudaSetDevice(ctx->device);
    int r = lb_twiddles_ensure(ctx, log_out);
    if (r) return r;
    const size_t n_in = (size_t)1 << log_in, n_out = (size_t)1 << log_out;
    if (chunk_cols <= 0) {
        // ~32 MiB of output per chunk: long enough copies for full PCIe rate, short enough pipeline fill
        chunk_cols = (int)(((size_t)8 << 20) / n_out);
        if (chunk_cols < 1) chunk_cols = 1;
    }
    if (chunk_cols > n_cols) chunk_cols = n_cols;
    const bool in_place = (log_in == log_out);
    // staging: each slot holds one chunk of inputs and (unless in place) one chunk of outputs
    size_t slot_words = (size_t)chunk_cols * (in_place ? n_in : n_in + n_out);
    if (!ctx->s_in) {
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking), "lde_host/stream");
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking), "lde_host/stream");
        for (int i = 0; i < lb_ctx::PIPE_SLOTS; ++i) {
            CK(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming), "lde_host/event");
            CK(cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming), "lde_host/event");
            CK(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming), "lde_host/event");
        }
    }
    if (slot_words > ctx->pipe_buf_words) {
        CK(cudaDeviceSynchronize(), "lde_host/sync");
        for (int i = 0; i < lb_ctx::PIPE_SLOTS; ++i) {
            if (ctx->pipe_buf[i]) cudaFree(ctx->pipe_buf[i]);
            ctx->pipe_buf[i] = nullptr;
        }
        ctx->pipe_buf_words = 0;
        for (int i = 0; i < lb_ctx::PIPE_SLOTS; ++i) CK(cudaMalloc(&ctx->pipe_buf[i], slot_words * sizeof(uint32_t)), "lde_host/alloc");
        ctx->pipe_buf_words = slot_words;
    }
    // the compute stream may still be running earlier work that owns the twiddle tables etc.: order after it
    int n_chunks = (n_cols + chunk_cols - 1) / chunk_cols;
    for (int k = 0; k < n_chunks; ++k) {
        int slot = k % lb_ctx::PIPE_SLOTS;
        int c0 = k * chunk_cols, nc = (c0 + chunk_cols <= n_cols) ? chunk_cols : (n_cols - c0);
        uint32_t* d_in = ctx->pipe_buf[slot];
        uint32_t* d_out = in_place ? d_in : d_in + (size_t)chunk_cols * n_in;
        if (k >= lb_ctx::PIPE_SLOTS) CK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_out[slot], 0), "lde_host/wait");  // slot drained
        CK(cudaMemcpyAsync(d_in, h_values + (size_t)c0 * n_in, (size_t)nc * n_in * 4, cudaMemcpyHostToDevice, ctx->s_in), "lde_host/h2d");
        CK(cudaEventRecord(ctx->ev_in[slot], ctx->s_in), "lde_host/record");
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[slot], 0), "lde_host/wait");
        CK(lb::cfft_interpolate(&ctx->tw, d_in, n_in, nc, log_in, ctx->sm_count, ctx->stream), "lde_host/interpolate");
        if (h_coeffs) {
            // coefficients leave from the compute stream's view: record, copy on the output stream before the evaluate overwrites (in place)
            CK(cudaEventRecord(ctx->ev_comp[slot], ctx->stream), "lde_host/record");
            CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[slot], 0), "lde_host/wait");
            CK(cudaMemcpyAsync(h_coeffs + (size_t)c0 * n_in, d_in, (size_t)nc * n_in * 4, cudaMemcpyDeviceToHost, ctx->s_out), "lde_host/d2h");
            if (in_place) {
                CK(cudaEventRecord(ctx->ev_out[slot], ctx->s_out), "lde_host/record");
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[slot], 0), "lde_host/wait");
            }
        }
        CK(lb::cfft_evaluate(&ctx->tw, d_in, n_in, log_in, d_out, n_out, log_out, nc, ctx->sm_count, ctx->stream), "lde_host/evaluate");
        CK(cudaEventRecord(ctx->ev_comp[slot], ctx->stream), "lde_host/record");
        CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[slot], 0), "lde_host/wait");
        CK(cudaMemcpyAsync(h_evals + (size_t)c0 * n_out, d_out, (size_t)nc * n_out * 4, cudaMemcpyDeviceToHost, ctx->s_out), "lde_host/d2h");
        CK(cudaEventRecord(ctx->ev_out[slot], ctx->s_out), "lde_host/record");
    }
    CK(cudaStreamSynchronize(ctx->s_out), "lde_host/sync");
    CK(cudaStreamSynchronize(ctx->stream), "lde_host/sync");
    return LB_OK;
}

int lb_prove_with_lookups(lb_ctx* ctx, const lb_trace_table* tables, int n_tables, const lb_preprocessed_column* lut_columns,
                          int n_lut_columns, const lb_prove_config* cfg, uint8_t** proof_out, size_t* proof_len) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !proof_out || !proof_len) return fail(ctx, LB_ERR_BAD_ARG, "prove: bad args");
    *proof_out = nullptr;
    *proof_len = 0;
    std::vector<uint8_t> bytes;
    int r = lb::prove_impl(ctx, tables, n_tables, lut_columns, n_lut_columns, cfg, bytes);
    if (r) return r;
    uint8_t* p = (uint8_t*)std::malloc(bytes.size() ? bytes.size() : 1);
    if (!p) return fail(ctx, LB_ERR_OOM, "prove: host malloc");
    std::memcpy(p, bytes.data(), bytes.size());
    *proof_out = p;
    *proof_len = bytes.size();
    return LB_OK;
}

// ---- multi-GPU: one rank per GPU, NCCL over NVLink (SURVEY 8b lb_comm_init / 8e) --------------------------------------
int lb_comm_unique_id(uint8_t id_out[LB_COMM_ID_BYTES]) {
    if (!id_out) return LB_ERR_BAD_ARG;
    lb::NcclApi& api = lb::nccl_api();
    if (!api.load()) return LB_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == LB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != ncclSuccess) return LB_ERR_NCCL;
    std::memcpy(id_out, &id, sizeof(id));
    return LB_OK;
}

int lb_comm_init(lb_ctx* ctx, const uint8_t id[LB_COMM_ID_BYTES], int rank, int world, lb_comm** out) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !id || !out || world < 1 || world > 64 || rank < 0 || rank >= world || (world & (world - 1)))
        return fail(ctx, LB_ERR_BAD_ARG, "comm: bad args (the world size must be a power of two)");
    *out = nullptr;
    lb::NcclApi& api = lb::nccl_api();
    if (!api.load()) return fail(ctx, LB_ERR_NCCL, api.error.c_str());
    cudaSetDevice(ctx->device);
    ncclUniqueId nid;
    std::memcpy(&nid, id, sizeof(nid));
    lb_comm* c = new lb_comm();
    c->rank = rank;
    c->world = world;
    while ((1 << c->log_world) < world) ++c->log_world;
    c->ctx = ctx;
    ncclResult_t r = api.CommInitRank(&c->comm, world, nid, rank);
    if (r != ncclSuccess) {
        std::string msg = std::string("ncclCommInitRank: ") + api.GetErrorString(r);
        delete c;
        return fail(ctx, LB_ERR_NCCL, msg.c_str());
    }
    *out = c;
    return LB_OK;
}

void lb_comm_destroy(lb_comm* comm) {
    if (!comm) return;
    lb::DeviceGuard _dg(comm->ctx);
    if (comm->ctx) cudaStreamSynchronize(comm->ctx->stream);
    lb::release_symmetric(comm);  // peers' mappings first, then this rank's buffer (callers destroy collectively)
    if (comm->comm) lb::nccl_api().CommDestroy(comm->comm);
    delete comm;
}

int lb_comm_stats(const lb_comm* comm, int* rank, int* world, uint64_t* bytes_sent, uint64_t* bytes_received, int* n_collectives) {
    if (!comm) return LB_ERR_BAD_ARG;
    if (rank) *rank = comm->rank;
    if (world) *world = comm->world;
    if (bytes_sent) *bytes_sent = comm->bytes_sent + comm->bytes_peer_stored;  // NCCL sends + direct NVLink peer stores
    if (bytes_received) *bytes_received = comm->bytes_received;
    if (n_collectives) *n_collectives = comm->n_collectives;
    return LB_OK;
}

int lb_prove_sharded(lb_ctx* ctx, lb_comm* comm, const lb_trace_table* tables, int n_tables,
                     const lb_preprocessed_column* lut_columns, int n_lut_columns, const lb_prove_config* cfg,
                     uint8_t** proof_out, size_t* proof_len) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !comm || !proof_out || !proof_len) return fail(ctx, LB_ERR_BAD_ARG, "prove_sharded: bad args");
    *proof_out = nullptr;
    *proof_len = 0;
    std::vector<uint8_t> bytes;
    int r = lb::prove_impl(ctx, tables, n_tables, lut_columns, n_lut_columns, cfg, bytes, comm);
    if (r) return r;
    uint8_t* p = (uint8_t*)std::malloc(bytes.size() ? bytes.size() : 1);
    if (!p) return fail(ctx, LB_ERR_OOM, "prove: host malloc");
    std::memcpy(p, bytes.data(), bytes.size());
    *proof_out = p;
    *proof_len = bytes.size();
    return LB_OK;
}

int lb_prove(lb_ctx* ctx, const lb_trace_table* tables, int n_tables, const lb_prove_config* cfg, uint8_t** proof_out,
             size_t* proof_len) {
    lb::DeviceGuard _dg(ctx);
    return lb_prove_with_lookups(ctx, tables, n_tables, nullptr, 0, cfg, proof_out, proof_len);
}

void lb_free_host(void* p) { std::free(p); }

int lb_trace_inputs(lb_ctx* ctx, uint32_t node_id, const int32_t* d_vals, uint64_t n, uint32_t out_mult, uint32_t* d_rows,
                    uint64_t row0) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_vals || !d_rows || node_id >= lb::P || out_mult >= lb::P) return fail(ctx, LB_ERR_BAD_ARG, "trace_inputs: bad args");
    cudaSetDevice(ctx->device);
    lb::TraceOp p{};
    p.op = LB_OP_INPUTS;
    p.node_id = node_id;
    p.lhs = d_vals;
    p.n = n;
    p.group = 1;
    p.out_mult_all = out_mult;
    p.rows = d_rows + row0 * 7;
    CK(lb::trace_op(p, ctx->stream), "trace_inputs");
    return LB_OK;
}

static int trace_binary_api(lb_ctx* ctx, bool mul, uint32_t node_id, uint32_t lhs_id, uint32_t rhs_id, const int32_t* d_lhs,
                            const int32_t* d_rhs, uint64_t n, uint32_t out_mult, int32_t* d_out, uint32_t* d_rows, uint64_t row0) {
    if (!ctx || !d_lhs || !d_rhs || !d_out || !d_rows || node_id >= lb::P || lhs_id >= lb::P || rhs_id >= lb::P || out_mult >= lb::P)
        return fail(ctx, LB_ERR_BAD_ARG, "trace_add/mul: bad args");
    cudaSetDevice(ctx->device);
    lb::TraceOp p{};
    p.op = mul ? LB_OP_MUL : LB_OP_ADD;
    p.node_id = node_id; p.lhs_id = lhs_id; p.rhs_id = rhs_id;
    p.lhs = d_lhs; p.rhs = d_rhs;
    p.n = n;
    p.group = 1;
    p.out_mult_all = out_mult;
    p.out = d_out;
    p.rows = d_rows + row0 * (mul ? 16 : 15);
    CK(lb::trace_op(p, ctx->stream), "trace_binary");
    return LB_OK;
}

int lb_trace_add(lb_ctx* ctx, uint32_t node_id, uint32_t lhs_id, uint32_t rhs_id, const int32_t* d_lhs, const int32_t* d_rhs,
                 uint64_t n, uint32_t out_mult, int32_t* d_out, uint32_t* d_rows, uint64_t row0) {
    lb::DeviceGuard _dg(ctx);
    return trace_binary_api(ctx, false, node_id, lhs_id, rhs_id, d_lhs, d_rhs, n, out_mult, d_out, d_rows, row0);
}

int lb_trace_mul(lb_ctx* ctx, uint32_t node_id, uint32_t lhs_id, uint32_t rhs_id, const int32_t* d_lhs, const int32_t* d_rhs,
                 uint64_t n, uint32_t out_mult, int32_t* d_out, uint32_t* d_rows, uint64_t row0) {
    lb::DeviceGuard _dg(ctx);
    return trace_binary_api(ctx, true, node_id, lhs_id, rhs_id, d_lhs, d_rhs, n, out_mult, d_out, d_rows, row0);
}

int lb_bit_reverse(lb_ctx* ctx, uint32_t* d_col, int log_size) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_col || log_size < 0 || log_size > 31) return fail(ctx, LB_ERR_BAD_ARG, "bit_reverse: bad args");
    cudaSetDevice(ctx->device);
    CK(lb::bit_reverse(d_col, log_size, ctx->stream), "bit_reverse");
    return LB_OK;
}

int lb_new_canonical_ordered(lb_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, int log_size) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_in || !d_out || d_in == d_out || log_size < 1 || log_size > 31)
        return fail(ctx, LB_ERR_BAD_ARG, "new_canonical_ordered: bad args");
    cudaSetDevice(ctx->device);
    CK(lb::canonical_to_storage(d_out, d_in, log_size, ctx->stream), "new_canonical_ordered");
    return LB_OK;
}

static int batch_inverse_api(lb_ctx* ctx, const uint32_t* const* in, uint32_t* const* out, int n_coords, size_t n) {
    cudaSetDevice(ctx->device);
    if (ensure_scratch(ctx, 64) != LB_OK) return LB_ERR_OOM;
    int* d_flag = (int*)ctx->d_scratch;
    CK(cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream), "batch_inverse: flag reset");
    if (n_coords == 1)
        CK(lb::batch_inverse_m31(out[0], in[0], n, d_flag, ctx->stream), "batch_inverse");
    else
        CK(lb::batch_inverse_qm31(out, in, n, d_flag, ctx->stream), "batch_inverse");
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "batch_inverse: flag read");
    CK(cudaStreamSynchronize(ctx->stream), "batch_inverse: sync");
    if (flag) return fail(ctx, LB_ERR_BAD_ARG, "0 has no inverse");
    return LB_OK;
}

int lb_batch_inverse_m31(lb_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, size_t n) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_in || !d_out) return fail(ctx, LB_ERR_BAD_ARG, "batch_inverse: bad args");
    return batch_inverse_api(ctx, &d_in, &d_out, 1, n);
}

int lb_batch_inverse_qm31(lb_ctx* ctx, const uint32_t* const d_in[4], uint32_t* const d_out[4], size_t n) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_in || !d_out) return fail(ctx, LB_ERR_BAD_ARG, "batch_inverse: bad args");
    for (int c = 0; c < 4; ++c)
        if (!d_in[c] || !d_out[c]) return fail(ctx, LB_ERR_BAD_ARG, "batch_inverse: null coordinate column");
    return batch_inverse_api(ctx, d_in, d_out, 4, n);
}

int lb_accumulate(lb_ctx* ctx, uint32_t* const d_column[4], const uint32_t* const d_other[4], size_t n) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_column || !d_other) return fail(ctx, LB_ERR_BAD_ARG, "accumulate: bad args");
    cudaSetDevice(ctx->device);
    for (int c = 0; c < 4; ++c) {
        if (!d_column[c] || !d_other[c]) return fail(ctx, LB_ERR_BAD_ARG, "accumulate: null coordinate column");
        CK(lb::add_inplace(d_column[c], d_other[c], n, ctx->stream), "accumulate");
    }
    return LB_OK;
}

int lb_generate_secure_powers(const uint32_t felt[4], int n_powers, uint32_t* h_out) {
    if (!felt || n_powers < 0 || (n_powers && !h_out)) return LB_ERR_BAD_ARG;
    for (int c = 0; c < 4; ++c)
        if (felt[c] >= lb::P) return LB_ERR_BAD_ARG;
    lb::QM31 f = lb::q_make(felt[0], felt[1], felt[2], felt[3]), acc = lb::q_from_m(1);
    for (int k = 0; k < n_powers; ++k) {
        h_out[4 * k] = acc.a.a; h_out[4 * k + 1] = acc.a.b; h_out[4 * k + 2] = acc.b.a; h_out[4 * k + 3] = acc.b.b;
        acc = lb::q_mul(acc, f);
    }
    return LB_OK;
}

int lb_trace_count_uses(lb_ctx* ctx, uint32_t* d_uses, const uint32_t* d_idx, uint64_t n_reads) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d_uses) return fail(ctx, LB_ERR_BAD_ARG, "trace_count_uses: bad args");
    cudaSetDevice(ctx->device);
    CK(lb::trace_count_uses(d_uses, d_idx, n_reads, ctx->stream), "trace_count_uses");
    return LB_OK;
}

int lb_trace_op(lb_ctx* ctx, const lb_trace_op_desc* d) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !d || !d->d_rows || !d->d_lhs || d->node_id >= lb::P || d->lhs_id >= lb::P || d->rhs_id >= lb::P)
        return fail(ctx, LB_ERR_BAD_ARG, "trace_op: bad args");
    int n_cols = 0;
    bool binary = false, lut = false, reduce = false;
    switch (d->op) {
        case LB_OP_ADD: n_cols = 15; binary = true; break;
        case LB_OP_MUL: case LB_OP_REM: n_cols = 16; binary = true; break;
        case LB_OP_LESS_THAN: n_cols = 22; binary = true; break;
        case LB_OP_RECIP: case LB_OP_SQRT: n_cols = 13; break;
        case LB_OP_SIN: case LB_OP_EXP2: case LB_OP_LOG2: n_cols = 12; lut = true; break;
        case LB_OP_SUM_REDUCE: n_cols = 14; reduce = true; break;
        case LB_OP_MAX_REDUCE: n_cols = 15; reduce = true; break;
        case LB_OP_INPUTS: n_cols = 7; break;
        case LB_OP_CONTIGUOUS: n_cols = 11; break;
        default: return fail(ctx, LB_ERR_BAD_ARG, "trace_op: unknown operator");
    }
    if (binary && !d->d_rhs) return fail(ctx, LB_ERR_BAD_ARG, "trace_op: binary operator without a right operand");
    if (d->op != LB_OP_INPUTS && !d->d_out) return fail(ctx, LB_ERR_BAD_ARG, "trace_op: no output tensor");
    if (reduce && d->group == 0) return fail(ctx, LB_ERR_BAD_ARG, "trace_op: reduction over an empty group");
    if ((lut || d->op == LB_OP_LESS_THAN) && (!d->lookup || !d->lookup->d_multiplicities))
        return fail(ctx, LB_ERR_BAD_ARG, "trace_op: lookup operator without its table");
    cudaSetDevice(ctx->device);
    lb::TraceOp p{};
    p.op = d->op;
    p.node_id = d->node_id; p.lhs_id = d->lhs_id; p.rhs_id = d->rhs_id;
    p.lhs = d->d_lhs; p.lhs_idx = d->d_lhs_idx; p.rhs = d->d_rhs; p.rhs_idx = d->d_rhs_idx;
    p.n = d->n;
    p.group = reduce ? d->group : 1;
    p.out_mult = d->d_out_mult;
    p.out = d->d_out;
    p.rows = d->d_rows + d->row0 * (uint64_t)n_cols;
    if (d->lookup) p.lut_mult = d->lookup->d_multiplicities;
    if (lut) {
        const lb_lookup& L = *d->lookup;
        if (L.n_ranges < 1 || L.n_ranges > LB_MAX_LOOKUP_RANGES || !L.d_values)
            return fail(ctx, LB_ERR_BAD_ARG, "trace_op: lookup layout must have 1..8 ranges and its value column");
        uint32_t base = 0;
        p.lut.n = L.n_ranges;
        for (int k = 0; k < L.n_ranges; ++k) {
            if (L.hi[k] < L.lo[k] || (k && L.lo[k] <= L.hi[k - 1]))
                return fail(ctx, LB_ERR_BAD_ARG, "trace_op: lookup ranges must be sorted and disjoint");
            p.lut.lo[k] = L.lo[k]; p.lut.hi[k] = L.hi[k]; p.lut.base[k] = base;
            base += (uint32_t)((int64_t)L.hi[k] - L.lo[k] + 1);
        }
        p.lut_vals = L.d_values;
        if (ensure_scratch(ctx, 64) != LB_OK) return LB_ERR_OOM;
        p.err = (int*)ctx->d_scratch;
        CK(cudaMemsetAsync(p.err, 0, sizeof(int), ctx->stream), "trace_op: flag reset");
    }
    CK(lb::trace_op(p, ctx->stream), "trace_op");
    if (lut) {  // the reference panics on an input outside the table; surface it here (one 4-byte read-back)
        int flag = 0;
        CK(cudaMemcpyAsync(&flag, p.err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "trace_op: flag read");
        CK(cudaStreamSynchronize(ctx->stream), "trace_op: sync");
        if (flag) return fail(ctx, LB_ERR_BAD_ARG, "Value should fit in range.");
    }
    return LB_OK;
}

int lb_prove_transcript(lb_ctx* ctx, uint8_t* out, size_t cap_hashes, size_t* n_hashes) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !n_hashes) return LB_ERR_BAD_ARG;
    *n_hashes = ctx->transcript.size();
    size_t n = ctx->transcript.size() < cap_hashes ? ctx->transcript.size() : cap_hashes;
    for (size_t i = 0; i < n && out; ++i) std::memcpy(out + 32 * i, ctx->transcript[i].b, 32);
    return LB_OK;
}

int lb_prove_stage_ms(lb_ctx* ctx, float* out, int cap, int* n) {
    lb::DeviceGuard _dg(ctx);
    if (!ctx || !n) return LB_ERR_BAD_ARG;
    *n = (int)ctx->stage_ms.size();
    for (int i = 0; i < cap && i < *n && out; ++i) out[i] = ctx->stage_ms[i];
    return LB_OK;
}

}  // extern "C"
