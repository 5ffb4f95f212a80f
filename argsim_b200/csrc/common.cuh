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

// ---- Philox4x32-10 counter RNG (Salmon et al. 2011), keyed by (seed, step | stream) ----------
// counter = (row_global, position, stream, 0).  Restated bit-exactly in numpy by
// argsim_b200/rng.py so the keep-mask / eps streams can be reproduced on the host.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}
// uniform in [0,1) with 24 bits
__host__ __device__ __forceinline__ float u01_24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
