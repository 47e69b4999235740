// rc_nccl.cpp — run-time binding of NCCL (see rc_nccl.h).
#include "rc_nccl.h"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

namespace rc {

const NcclApi* nccl_api(std::string& err) {
    static NcclApi api;
    static bool tried = false, ok = false;
    static std::string why;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        const char* env = std::getenv("RC_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* n : names) {
            if (!n || !*n) continue;
            // a library of that soname already mapped into the process (torch's bundled NCCL) is reused by the loader
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
            why = dlerror();
        }
        if (h) {
            auto sym = [&](const char* s) -> void* {
                void* p = dlsym(h, s);
                if (!p) why = std::string("missing symbol ") + s;
                return p;
            };
            api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
            api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
            api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
            api.GroupStart = (int (*)())sym("ncclGroupStart");
            api.GroupEnd = (int (*)())sym("ncclGroupEnd");
            api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
            api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
            ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd &&
                 api.GetErrorString;
        }
    }
    if (!ok) {
        err = "NCCL is not available (" + why + "): the multi-GPU path needs libnccl.so.2";
        return nullptr;
    }
    return &api;
}

}  // namespace rc
