// philox.h -- Philox4x32-10 counter RNG (Salmon et al. 2011), usable from host and device.
// Keyed by (seed_lo, step); counter = (row_global, position, stream, seed_hi).  Restated
// bit-exactly in numpy by argsim_b200/rng.py so keep-mask / eps streams can be reproduced.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define PHILOX_HD __host__ __device__ __forceinline__
#else
#define PHILOX_HD inline
#endif

PHILOX_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
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
PHILOX_HD float u01_24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// stream ids
enum { PHILOX_STREAM_KEEP = 0, PHILOX_STREAM_EPS = 1 };
