// NCCL is bound at run time (dlopen of the libnccl.so.2 that ships with torch) so that
// libbabelb200.so loads on hosts without NCCL or a GPU; only the slab halo exchange needs it.
#pragma once
#include <dlfcn.h>
#include <string>
#include <nccl.h>

struct NcclApi {
    bool ok = false;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static inline NcclApi &nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = nullptr;
    const char *env = getenv("BB_NCCL_LIB");
    const char *names[] = { env, "libnccl.so.2", "libnccl.so" };
    for (const char *n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { api.err = std::string("dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return api; }
#define BB_SYM(field, name)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));             \
    if (!api.field) { api.err = std::string("missing symbol ") + name; return api; }
    BB_SYM(GetUniqueId, "ncclGetUniqueId")
    BB_SYM(CommInitRank, "ncclCommInitRank")
    BB_SYM(CommDestroy, "ncclCommDestroy")
    BB_SYM(GroupStart, "ncclGroupStart")
    BB_SYM(GroupEnd, "ncclGroupEnd")
    BB_SYM(Send, "ncclSend")
    BB_SYM(Recv, "ncclRecv")
    BB_SYM(GetErrorString, "ncclGetErrorString")
#undef BB_SYM
    api.ok = true;
    return api;
}
