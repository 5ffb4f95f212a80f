// common.cuh -- shared helpers for the argsim_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

#define CUDA_CHECK(x)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (x);                                                                \
        if (e_ != cudaSuccess)                                                               \
            throw std::runtime_error(std::string(#x) + " failed at " + __FILE__ + ":" +      \
                                     std::to_string(__LINE__) + ": " + cudaGetErrorString(e_)); \
    } while (0)

typedef __nv_bfloat16 bf16;

// Launch counter (bench.py reports it as gpu_launches).
extern long long g_launch_count;
#define COUNT_LAUNCH() (++g_launch_count)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// A dense row-major matrix living in HBM, optionally with a bf16 twin used as tensor-core operand.
struct Mat {
    float* f = nullptr;   // fp32 view (may be null when only the bf16 twin exists)
    bf16* h = nullptr;    // bf16 twin (null in fp32-validate mode)
    long long rows = 0;
    int cols = 0;
    int ld = 0;           // leading dimension in elements (same for both views)
    Mat() {}
    Mat(float* f_, bf16* h_, long long r, int c, int ld_) : f(f_), h(h_), rows(r), cols(c), ld(ld_) {}
    Mat colslice(int c0, int n) const { return Mat(f ? f + c0 : nullptr, h ? h + c0 : nullptr, rows, n, ld); }
    Mat rowslice(long long r0, long long n) const {
        return Mat(f ? f + r0 * ld : nullptr, h ? h + r0 * ld : nullptr, n, cols, ld);
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

#include "philox.h"
