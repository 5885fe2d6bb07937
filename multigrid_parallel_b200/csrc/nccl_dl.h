// nccl_dl.h -- NCCL bound at run time (dlopen), so libmgb.so has no link-time
// NCCL dependency: a single-GPU process never loads it, and a multi-rank
// process shares whichever libnccl.so.2 is already mapped (e.g. torch's).
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // types and enums only

namespace mgb {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *error = nullptr;

    bool load()
    {
        if (handle)
            return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle)
                break;
        }
        if (!handle) {
            error = "cannot dlopen libnccl.so.2";
            return false;
        }
#define MGB_SYM(field, name)                                            \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));     \
    if (!field) {                                                       \
        error = "missing NCCL symbol " name;                            \
        return false;                                                   \
    }
        MGB_SYM(GetUniqueId, "ncclGetUniqueId")
        MGB_SYM(CommInitRank, "ncclCommInitRank")
        MGB_SYM(CommDestroy, "ncclCommDestroy")
        MGB_SYM(GetErrorString, "ncclGetErrorString")
        MGB_SYM(GroupStart, "ncclGroupStart")
        MGB_SYM(GroupEnd, "ncclGroupEnd")
        MGB_SYM(Send, "ncclSend")
        MGB_SYM(Recv, "ncclRecv")
        MGB_SYM(AllReduce, "ncclAllReduce")
        MGB_SYM(Broadcast, "ncclBroadcast")
        MGB_SYM(AllGather, "ncclAllGather")
#undef MGB_SYM
        return true;
    }
};

inline NcclApi &nccl()
{
    static NcclApi api;
    return api;
}

}  // namespace mgb
