// nccl_dyn.h -- NCCL bound at run time (dlopen): the library has no link-time dependency on it, and inside a process that
// already loaded a libnccl (torch) that one is used.  Definition in dist_multi.cu.
#pragma once
#include <nccl.h>
#include "common.h"

namespace fpm {

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
};

NcclApi* nccl_api();     // nullptr (and fpm_last_error set) when NCCL cannot be loaded

#define FPM_NCCL(api, call)                                                                                          \
    do {                                                                                                             \
        ncclResult_t r__ = (call);                                                                                   \
        if (r__ != ncclSuccess) { fpm::set_error("NCCL error %d (%s): %s", (int)r__, (api)->GetErrorString(r__), #call); return FPM_ERR_COMM; } \
    } while (0)

}  // namespace fpm
