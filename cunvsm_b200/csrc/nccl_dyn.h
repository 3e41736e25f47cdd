// NCCL bound at run time (dlopen) so that libnvsm_b200.so has no link-time dependency on
// it: single-GPU users never load NCCL, and inside a torch process the already-loaded
// bundled libnccl.so.2 is reused. Only the six entry points the step needs are resolved.
#pragma once

#include <dlfcn.h>
#include <stddef.h>

namespace nvsm {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { kNcclInt8 = 0, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, void* /*cudaStream_t*/) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, void* /*cudaStream_t*/) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    bool load(const char** why) {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { *why = "dlopen(libnccl.so.2) failed"; return false; }
        GetUniqueId = (int (*)(NcclUniqueId*))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(handle, "ncclCommInitRank");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, void*))dlsym(handle, "ncclAllReduce");
        AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, void*))dlsym(handle, "ncclAllGather");
        CommDestroy = (int (*)(NcclComm))dlsym(handle, "ncclCommDestroy");
        GetErrorString = (const char* (*)(int))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !AllGather || !CommDestroy || !GetErrorString) {
            *why = "libnccl is missing a required symbol";
            return false;
        }
        return true;
    }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace nvsm
