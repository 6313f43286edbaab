// gem_b200/csrc/nccl_api.h -- NCCL entry points resolved with dlsym (types from <nccl.h>).
#pragma once
#include <nccl.h>
namespace gemb {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
};
}  // namespace gemb
