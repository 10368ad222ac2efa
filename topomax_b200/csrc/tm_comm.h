// NCCL plumbing for the row-strip sharded solver (one process per GPU).
//
// The library has no link-time dependency on NCCL: libnccl.so.2 is resolved with dlopen at
// the first multi-rank call (in a torch process that is the copy torch.distributed already
// loaded), so the same .so loads on machines without NCCL for the single-GPU path.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <string>

namespace tmx {

// minimal ABI subset of nccl.h (stable since NCCL 2.x)
struct NcclUniqueId {
    char internal[128];
};
using NcclComm = void*;
enum { kNcclSum = 0 };
enum { kNcclInt8 = 0, kNcclFloat32 = 7, kNcclFloat64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool loaded = false;
    std::string error;

    bool load() {
        if (loaded) return true;
        void* h = nullptr;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            error = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return false;
        }
#define TM_NCCL_SYM(field, sym)                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(h, sym));          \
    if (!field) {                                                      \
        error = std::string("libnccl is missing symbol ") + sym;       \
        return false;                                                  \
    }
        TM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        TM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        TM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        TM_NCCL_SYM(AllReduce, "ncclAllReduce")
        TM_NCCL_SYM(Broadcast, "ncclBroadcast")
        TM_NCCL_SYM(Send, "ncclSend")
        TM_NCCL_SYM(Recv, "ncclRecv")
        TM_NCCL_SYM(GroupStart, "ncclGroupStart")
        TM_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        TM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef TM_NCCL_SYM
        loaded = true;
        return true;
    }
};

inline NcclApi& nccl() {
    static NcclApi api;
    return api;
}

}  // namespace tmx
