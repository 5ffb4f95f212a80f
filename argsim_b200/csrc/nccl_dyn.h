// nccl_dyn.h -- the handful of NCCL entry points the data-parallel path uses, resolved with
// dlopen at first use so that the library loads (and single-GPU runs work) without libnccl.
// When torch is already imported its bundled libnccl.so.2 is the one the loader hands back.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <stdexcept>
#include <string>

struct NcclApi {
    typedef struct ncclComm* comm_t;
    typedef struct { char internal[128]; } unique_id;
    enum { Sum = 0 };
    enum { Float32 = 7, Float64 = 8 };  // ncclDataType_t values (nccl.h)
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* lib = nullptr;

    static NcclApi& get() {
        static NcclApi api;
        if (!api.lib) api.load();
        return api;
    }
    void load() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) throw std::runtime_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString)
            throw std::runtime_error("libnccl is missing a required symbol");
    }
    void check(int rc, const char* what) {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + GetErrorString(rc));
    }
};
