// comm.hpp -- NCCL inside the library (SURVEY.md 8b "Threading", 8e).
//
// The one exchange of the path is the sum over column shards of the per-pair partial core counts
// (all-reduce for sampled pairs, reduce-scatter for all-pairs row blocks). NCCL is bound at run time
// with dlopen("libnccl.so.2"): a process that already carries an NCCL (torch bundles its own) keeps
// that single instance, a plain C/C++/Rust host picks up the system library, and a single-GPU host
// needs none at all. Only the handful of entry points below is used; the types come from <nccl.h>.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace pansim {

struct NcclApi {
    void *handle = nullptr;
    std::string error;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;

    bool ok() const { return handle != nullptr; }
};

// process-wide, loaded on first use; nullptr-handle + error text when no NCCL can be found
inline NcclApi &nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
        return api;
    }
    bool all = true;
    auto sym = [&](const char *name) -> void * {
        void *p = dlsym(api.handle, name);
        if (!p) { all = false; api.error = std::string("NCCL symbol missing: ") + name; }
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    if (!all) {
        dlclose(api.handle);
        api.handle = nullptr;
    }
    return api;
}

}  // namespace pansim
