// CPU prover (oracle; TEST INFRASTRUCTURE ONLY).
//
// Column buffers: uninitialised on allocation (the first touch happens in the parallel loop that fills them) and recycled
// through a size-keyed free list, so repeated proofs do not pay the kernel's page-fault path for every multi-megabyte column
// (glibc would mmap / munmap each of them).
#pragma once
#include <omp.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace cpu {

struct Pool {
    std::mutex mu;
    std::unordered_map<size_t, std::vector<void*>> free_list;
    static Pool& get() {
        static Pool* p = new Pool();  // never destroyed: buffers may outlive static destruction order
        return *p;
    }
    void* alloc(size_t bytes) {
        bytes = (bytes + 63) & ~(size_t)63;
        if (bytes == 0) bytes = 64;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_list.find(bytes);
            if (it != free_list.end() && !it->second.empty()) {
                void* p = it->second.back();
                it->second.pop_back();
                return p;
            }
        }
        void* p = aligned_alloc(64, bytes);
        if (!p) throw std::bad_alloc();
        return p;
    }
    void release(void* p, size_t bytes) {
        bytes = (bytes + 63) & ~(size_t)63;
        if (bytes == 0) bytes = 64;
        if (bytes < (1u << 16)) {  // small blocks go back to malloc
            free(p);
            return;
        }
        std::lock_guard<std::mutex> g(mu);
        free_list[bytes].push_back(p);
    }
    void trim() {
        std::lock_guard<std::mutex> g(mu);
        for (auto& kv : free_list)
            for (void* p : kv.second) free(p);
        free_list.clear();
    }
};

static inline void par_fill(uint32_t* p, size_t n, uint32_t v) {
#pragma omp parallel for schedule(static) if (n >= (1u << 16))
    for (size_t i = 0; i < n; i++) p[i] = v;
}
static inline void par_copy(uint32_t* dst, const uint32_t* src, size_t n) {
    if (n < (1u << 16)) {
        if (n) memcpy(dst, src, n * 4);
        return;
    }
    const size_t CH = 1u << 14;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i += CH) memcpy(dst + i, src + i, (n - i < CH ? n - i : CH) * 4);
}

// the subset of std::vector<uint32_t> the prover uses
struct Col {
    uint32_t* p = nullptr;
    size_t n = 0;
    Col() {}
    Col(size_t n_, uint32_t v) { assign(n_, v); }
    Col(const uint32_t* b, const uint32_t* e) {
        alloc_uninit(e - b);
        par_copy(p, b, n);
    }
    Col(const Col& o) {
        alloc_uninit(o.n);
        par_copy(p, o.p, n);
    }
    Col(Col&& o) noexcept : p(o.p), n(o.n) {
        o.p = nullptr;
        o.n = 0;
    }
    Col& operator=(const Col& o) {
        if (this != &o) {
            alloc_uninit(o.n);
            par_copy(p, o.p, n);
        }
        return *this;
    }
    Col& operator=(Col&& o) noexcept {
        if (this != &o) {
            reset();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~Col() { reset(); }
    void reset() {
        if (p) Pool::get().release(p, n * 4);
        p = nullptr;
        n = 0;
    }
    void alloc_uninit(size_t n_) {
        reset();
        n = n_;
        p = (uint32_t*)Pool::get().alloc(n * 4);
    }
    void assign(size_t n_, uint32_t v) {
        alloc_uninit(n_);
        par_fill(p, n, v);
    }
    uint32_t* data() { return p; }
    const uint32_t* data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    uint32_t& operator[](size_t i) { return p[i]; }
    const uint32_t& operator[](size_t i) const { return p[i]; }
};

}  // namespace cpu
