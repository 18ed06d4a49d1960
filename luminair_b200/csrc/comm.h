// Multi-GPU plumbing of the sharded prover: one process (or thread) per GPU, NCCL over NVLink / NVSwitch for the exchanges.
// NCCL is bound at run time (dlopen of libnccl.so.2 - inside a PyTorch process that resolves to the copy torch already
// loaded), so the library keeps working on a single GPU with no NCCL installed.  Only the types come from <nccl.h>.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>
#include <vector>

#include "ctx.h"

namespace lb {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = "NCCL not found (dlopen libnccl.so.2)";
            return false;
        }
#define LB_NCCL_SYM(field, name)                                    \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name)); \
    if (!field) {                                                   \
        error = std::string("NCCL symbol missing: ") + name;        \
        return false;                                               \
    }
        LB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        LB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        LB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        LB_NCCL_SYM(GroupStart, "ncclGroupStart")
        LB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        LB_NCCL_SYM(Send, "ncclSend")
        LB_NCCL_SYM(Recv, "ncclRecv")
        LB_NCCL_SYM(AllGather, "ncclAllGather")
        LB_NCCL_SYM(AllReduce, "ncclAllReduce")
        LB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef LB_NCCL_SYM
        return true;
    }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace lb

struct lb_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, log_world = 0;
    lb_ctx* ctx = nullptr;
    // traffic of the last lb_prove_sharded on this rank
    unsigned long long bytes_sent = 0, bytes_received = 0;
    int n_collectives = 0;
    // Symmetric heap for the fused exchange: every rank allocates the same-sized buffer and maps every peer's copy (CUDA IPC,
    // NVLink peer access), so a kernel on rank a can store straight into rank b's buffer at the same offset.  sym_peer[r] is
    // this process's mapping of rank r's buffer (sym_peer[rank] == sym_base).  ipc: 1 usable, 0 unavailable (same-process
    // ranks, IPC refused, LB_SHARD_IPC=0): the exchange then goes through NCCL send/recv.
    uint32_t* sym_base = nullptr;
    size_t sym_words = 0, sym_used = 0;
    std::vector<uint32_t*> sym_peer;
    int ipc = -1;
    unsigned long long bytes_peer_stored = 0;  // written into peers' memory by the fused LDE / shift kernels
};

namespace lb {
inline void release_symmetric(lb_comm* c) {
    for (int r = 0; r < (int)c->sym_peer.size(); ++r)
        if (r != c->rank && c->sym_peer[r]) cudaIpcCloseMemHandle(c->sym_peer[r]);
    c->sym_peer.clear();
    if (c->sym_base) cudaFree(c->sym_base);
    c->sym_base = nullptr;
    c->sym_words = 0;
}
}  // namespace lb
