// NCCL plumbing behind the C ABI ("one ncclAllGather of per-shard top-k").  NCCL is resolved at
// run time with dlopen -- the library carries no link dependency on it, a single-GPU host never loads it, and inside a
// Python process that already holds torch's bundled libnccl.so.2 the same copy is reused (same soname).
#pragma once
#include "common.cuh"

#include <nccl.h>  // types and enums only

namespace cb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

// NULL (with cb_last_error set) when libnccl cannot be loaded
const NcclApi* nccl_api();

}  // namespace cb

struct cb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};
