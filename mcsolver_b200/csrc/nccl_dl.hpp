// NCCL through dlopen (like NVRTC): no link-time dependency, no PyTorch.  Shared by the tempering driver (pt.cu) and the slab
// decomposition (slab.cu).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

#include "common.cuh"

namespace mcg {

// ---- NCCL through dlopen ----
struct NcclApi {
    bool ok = false;
    std::string why;
    decltype(&ncclGetUniqueId) getUniqueId;
    decltype(&ncclCommInitRank) commInitRank;
    decltype(&ncclCommDestroy) commDestroy;
    decltype(&ncclAllGather) allGather;
    decltype(&ncclAllReduce) allReduce;
    decltype(&ncclSend) send;
    decltype(&ncclRecv) recv;
    decltype(&ncclGroupStart) groupStart;
    decltype(&ncclGroupEnd) groupEnd;
    decltype(&ncclGetErrorString) getErrorString;
};
inline NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = nullptr;
        const char *env = getenv("MCG_NCCL_LIB");
        if (env && env[0]) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        for (const char *n : {"libnccl.so.2", "libnccl.so"})
            if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.why = "cannot dlopen libnccl.so.2 (set MCG_NCCL_LIB to its path)"; return; }
#define MCG_SYM(field, name)                                                  \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));        \
    if (!api.field) { api.why = std::string("libnccl lacks ") + name; return; }
        MCG_SYM(getUniqueId, "ncclGetUniqueId")
        MCG_SYM(commInitRank, "ncclCommInitRank")
        MCG_SYM(commDestroy, "ncclCommDestroy")
        MCG_SYM(allGather, "ncclAllGather")
        MCG_SYM(allReduce, "ncclAllReduce")
        MCG_SYM(send, "ncclSend")
        MCG_SYM(recv, "ncclRecv")
        MCG_SYM(groupStart, "ncclGroupStart")
        MCG_SYM(groupEnd, "ncclGroupEnd")
        MCG_SYM(getErrorString, "ncclGetErrorString")
#undef MCG_SYM
        api.ok = true;
    });
    return api;
}
#define MCG_NCCL(call)                                                                                                     \
    do {                                                                                                                   \
        ncclResult_t r_ = (call);                                                                                          \
        if (r_ != ncclSuccess) throw ::mcg::Error(MCG_ERR_NCCL, std::string(#call) + ": " + nccl_api().getErrorString(r_)); \
    } while (0)


}  // namespace mcg
