// extern "C" boundary of libluminair_b200.so — see include/luminair_b200.h.
#include "../../include/luminair_b200.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "cfft.cuh"
#include "merkle.cuh"

struct lb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    size_t total_mem = 0;
    lb::Twiddles tw;
    std::string err;
    // scratch for pointer tables / index lists
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
};

namespace {
int fail(lb_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        ctx->err = what;
        if (e != cudaSuccess) {
            ctx->err += ": ";
            ctx->err += cudaGetErrorString(e);
        }
    }
    return code;
}
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

int lb_version(void) { return 1; }

int lb_ctx_create(int device, lb_ctx** out) {
    if (!out) return LB_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return LB_ERR_CUDA;
    lb_ctx* ctx = new lb_ctx();
    ctx->device = device;
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
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = prop.totalGlobalMem;
    *out = ctx;
    return LB_OK;
}

void lb_ctx_destroy(lb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    lb::twiddles_destroy(&ctx->tw);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* lb_last_error(lb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int lb_sync(lb_ctx* ctx) {
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    return LB_OK;
}

int lb_alloc(lb_ctx* ctx, size_t n_u32, uint32_t** d_out) {
    if (!ctx || !d_out) return LB_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    void* p = nullptr;
    CK(cudaMalloc(&p, (n_u32 ? n_u32 : 1) * sizeof(uint32_t)), "alloc");
    *d_out = (uint32_t*)p;
    return LB_OK;
}

int lb_free(lb_ctx* ctx, uint32_t* d_ptr) {
    if (!ctx) return LB_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream), "free/sync");
    CK(cudaFree(d_ptr), "free");
    return LB_OK;
}

int lb_memset_zero(lb_ctx* ctx, uint32_t* d_ptr, size_t n_u32) {
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaMemsetAsync(d_ptr, 0, n_u32 * sizeof(uint32_t), ctx->stream), "memset");
    return LB_OK;
}

int lb_upload(lb_ctx* ctx, uint32_t* d_dst, const uint32_t* h_src, size_t n_u32) {
    if (!ctx || (!h_src && n_u32)) return LB_ERR_BAD_ARG;
    CK(cudaMemcpyAsync(d_dst, h_src, n_u32 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream), "upload");
    CK(cudaStreamSynchronize(ctx->stream), "upload/sync");
    return LB_OK;
}

int lb_download(lb_ctx* ctx, uint32_t* h_dst, const uint32_t* d_src, size_t n_u32) {
    if (!ctx || (!h_dst && n_u32)) return LB_ERR_BAD_ARG;
    CK(cudaMemcpyAsync(h_dst, d_src, n_u32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream), "download");
    CK(cudaStreamSynchronize(ctx->stream), "download/sync");
    return LB_OK;
}

int lb_copy(lb_ctx* ctx, uint32_t* d_dst, const uint32_t* d_src, size_t n_u32) {
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaMemcpyAsync(d_dst, d_src, n_u32 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream), "copy");
    return LB_OK;
}

int lb_timer_start(lb_ctx* ctx) {
    if (!ctx) return LB_ERR_BAD_ARG;
    CK(cudaEventRecord(ctx->ev0, ctx->stream), "timer start");
    return LB_OK;
}

int lb_timer_stop_ms(lb_ctx* ctx, float* ms_out) {
    if (!ctx || !ms_out) return LB_ERR_BAD_ARG;
    CK(cudaEventRecord(ctx->ev1, ctx->stream), "timer stop");
    CK(cudaEventSynchronize(ctx->ev1), "timer sync");
    CK(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1), "timer elapsed");
    return LB_OK;
}

int lb_device_info(lb_ctx* ctx, int* sm_count, size_t* total_mem_bytes) {
    if (!ctx) return LB_ERR_BAD_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (total_mem_bytes) *total_mem_bytes = ctx->total_mem;
    return LB_OK;
}

int lb_twiddles_ensure(lb_ctx* ctx, int max_log) {
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
    if (!ctx || !d_out) return LB_ERR_BAD_ARG;
    if (ctx->tw.max_log < root_log + 1) return fail(ctx, LB_ERR_BAD_ARG, "twiddles_export: tables too small");
    CK(lb::twiddles_export_stwo(&ctx->tw, root_log, d_out, ctx->stream), "twiddles export");
    return LB_OK;
}

int lb_interpolate_batch(lb_ctx* ctx, uint32_t* d_cols, size_t stride, int n_cols, int log_size) {
    if (!ctx || n_cols < 0 || log_size < 1) return fail(ctx, LB_ERR_BAD_ARG, "interpolate: bad args");
    if (stride < ((size_t)1 << log_size) && n_cols > 1) return fail(ctx, LB_ERR_BAD_ARG, "interpolate: stride < column size");
    int r = lb_twiddles_ensure(ctx, log_size);
    if (r) return r;
    CK(lb::cfft_interpolate(&ctx->tw, d_cols, stride, n_cols, log_size, ctx->sm_count, ctx->stream), "interpolate");
    return LB_OK;
}

int lb_evaluate_batch(lb_ctx* ctx, const uint32_t* d_coeffs, size_t src_stride, int log_in, uint32_t* d_out,
                      size_t dst_stride, int log_out, int n_cols) {
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
    if (!ctx || n_cols < 0 || n_idx < 0) return fail(ctx, LB_ERR_BAD_ARG, "gather: bad args");
    if (n_cols == 0 || n_idx == 0) return LB_OK;
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

}  // extern "C"
