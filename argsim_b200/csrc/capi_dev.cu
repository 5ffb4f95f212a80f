// capi_dev.cu -- C ABI of libargsim_b200_dev.so (include/argsim_b200_dev.h): development microbenchmarks, kept out of the
// product library.
#include "../../include/argsim_b200_dev.h"
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

int xbench_run(int device, int method, int groups, int rows, int iters, double* cycles_per_iter, int* max_clusters);

static std::string g_dev_err;

extern "C" {
int argsim_bench_exchange(int32_t device, int32_t method, int32_t groups, int32_t rows, int32_t iters, double* cycles_per_round,
                          int32_t* max_clusters) {
    try {
        int mc = 0;
        xbench_run(device, method, groups, rows, iters, cycles_per_round, &mc);
        if (max_clusters) *max_clusters = mc;
    } catch (const std::exception& ex) {
        g_dev_err = ex.what();
        cudaGetLastError();
        return -2;
    }
    return 0;
}
const char* argsim_dev_last_error(void) { return g_dev_err.c_str(); }
}
