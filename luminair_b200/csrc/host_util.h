// Host-side utilities shared by prover.cu and capi.cu: error type, stream-ordered arena.
#pragma once
#include <algorithm>
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
    Arena(cudaMemPool_t pool, cudaStream_t s) : pool_(pool), stream_(s) {}
    ~Arena() {
        for (void* p : ptrs_) cudaFreeAsync(p, stream_);
    }
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
    template <class T>
    T* upload(const std::vector<T>& v) {
        T* d = alloc<T>(v.size());
        if (!v.empty()) ck(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream_), "upload");
        // pageable source: make sure the staging copy is complete before `v` can go away
        ck(cudaStreamSynchronize(stream_), "upload/sync");
        return d;
    }

   private:
    cudaMemPool_t pool_;
    cudaStream_t stream_;
    std::vector<void*> ptrs_;
};


}  // namespace lb
