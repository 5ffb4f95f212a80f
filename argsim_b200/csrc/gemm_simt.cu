// gemm_simt.cu -- fp32 FMA GEMM used by the FP32_VALIDATE precision mode (parity runs against
// the fp64/fp32 oracle at 1e-3) and as the on-device checker of the tcgen05 path in tests.
// Plain 64x64x16 shared-memory tiling, 4x4 register micro-tile; not a performance path.
#include "kernels.h"

namespace {
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) k_gemm_simt(const float* __restrict__ A, int lda, int a_mn,
                                                   const float* __restrict__ B, int ldb, int b_mn,
                                                   float* __restrict__ C, int ldc, int M, int N, int K, float alpha,
                                                   const float* __restrict__ bias, int accumulate, bf16* __restrict__ Ch) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int idx = tid + it * 256;
            int m, k;
            if (a_mn) { k = idx >> 6; m = idx & 63; } else { m = idx >> 4; k = idx & 15; }
            float v = 0.f;
            if (m0 + m < M && k0 + k < K)
                v = a_mn ? A[(long long)(k0 + k) * lda + m0 + m] : A[(long long)(m0 + m) * lda + k0 + k];
            As[k][m] = v;
            int n;
            if (b_mn) { k = idx >> 6; n = idx & 63; } else { n = idx >> 4; k = idx & 15; }
            v = 0.f;
            if (n0 + n < N && k0 + k < K)
                v = b_mn ? B[(long long)(k0 + k) * ldb + n0 + n] : B[(long long)(n0 + n) * ldb + k0 + k];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = alpha * acc[i][j];
            if (bias) v += bias[n];
            long long o = (long long)m * ldc + n;
            if (C) {
                if (accumulate) v += C[o];
                C[o] = v;
            }
            if (Ch) Ch[o] = __float2bfloat16(v);
        }
    }
}
}  // namespace

void gemm_simt(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, float* C, int ldc, int M, int N,
               int K, float alpha, const float* bias, int accumulate, bf16* C_h, cudaStream_t s) {
    if (M <= 0 || N <= 0) return;
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    k_gemm_simt<<<grid, 256, 0, s>>>(A, lda, a_mn, B, ldb, b_mn, C, ldc, M, N, K, alpha, bias, accumulate, C_h);
    COUNT_LAUNCH();
}
