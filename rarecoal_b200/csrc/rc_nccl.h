// rc_nccl.h — NCCL bound at run time (dlopen), so that librarecoal_b200.so has no link-time dependency on it: a
// single-GPU caller never touches NCCL, a multi-GPU caller gets the libnccl.so.2 already loaded in its process (e.g.
// torch's) or the system's.  There is no substitute path: if NCCL cannot be loaded the multi-GPU entry points fail.
//
// Only the stable C ABI of NCCL 2.x is used; the few types needed are restated here (nccl.h: ncclUniqueId is 128 opaque
// bytes passed by value, ncclFloat64 = 8, ncclSum = 0).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <string>

namespace rc {

struct NcclUniqueId {
    char internal[128];
};
typedef struct ncclComm* NcclComm;

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
enum { kNcclFloat64 = 8, kNcclSum = 0 };

// The process-wide binding; nullptr (with a message in err) when libnccl cannot be loaded.  RC_NCCL_LIB overrides the name.
const NcclApi* nccl_api(std::string& err);

}  // namespace rc
