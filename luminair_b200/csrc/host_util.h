// Host-side utilities shared by prover.cu and capi.cu: error type, stream-ordered arena.
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.h"

namespace lb {

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
struct ProveError {
    int code;
    std::string msg;
};
[[noreturn]] inline void fail(int code, const std::string& msg) { throw ProveError{code, msg}; }
inline void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess)
        fail(e == cudaErrorMemoryAllocation ? LB_ERR_OOM : LB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------
// device memory for one prove() call: stream-ordered allocations, released together
// ------------------------------------------------------------------------------------
class Arena {
   public:
    explicit Arena(lb_ctx* ctx) : ctx_(ctx), pool_(ctx->pool), stream_(ctx->stream) {
        // the staging buffer is ours once the previous user's copies have left it
        if (ctx_->stage_in_flight) {
            if (cudaEventQuery(ctx_->ev_stage) != cudaSuccess) cudaEventSynchronize(ctx_->ev_stage);
            cudaGetLastError();
            ctx_->stage_in_flight = false;
        }
    }
    ~Arena() {
        for (void* p : ptrs_) cudaFreeAsync(p, stream_);
        if (staged_ && cudaEventRecord(ctx_->ev_stage, stream_) == cudaSuccess) ctx_->stage_in_flight = true;
    }
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;
    template <class T>
    T* alloc(size_t n) {
        void* p = nullptr;
        ck(cudaMallocFromPoolAsync(&p, (n ? n : 1) * sizeof(T), pool_, stream_), "device alloc");
        ptrs_.push_back(p);
        return (T*)p;
    }
    void release(void* p) {
        auto it = std::find(ptrs_.begin(), ptrs_.end(), p);
        if (it != ptrs_.end()) {
            cudaFreeAsync(p, stream_);
            ptrs_.erase(it);
        }
    }
    // host -> device, stream-ordered; `h` may be reused as soon as the call returns
    void upload_to(void* d, const void* h, size_t bytes) {
        if (!bytes) return;
        const size_t padded = (bytes + 15) & ~(size_t)15;
        if (ctx_->h_stage && staged_ + padded <= ctx_->h_stage_bytes) {
            uint8_t* s = ctx_->h_stage + staged_;
            std::memcpy(s, h, bytes);
            staged_ += padded;
            ck(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, stream_), "upload");
            return;
        }
        // pageable source: make sure the staging copy is complete before `h` can go away
        ck(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream_), "upload");
        ck(cudaStreamSynchronize(stream_), "upload/sync");
    }
    template <class T>
    T* upload(const std::vector<T>& v) {
        T* d = alloc<T>(v.size());
        upload_to(d, v.data(), v.size() * sizeof(T));
        return d;
    }

   private:
    lb_ctx* ctx_;
    cudaMemPool_t pool_;
    cudaStream_t stream_;
    std::vector<void*> ptrs_;
    size_t staged_ = 0;
};


}  // namespace lb
