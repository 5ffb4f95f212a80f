// kernels.cu -- HBM-bound kernels of the VAE step: embedding gather / scatter-add, latent
// (reparameterisation + KL), fused softmax cross-entropy with gradient, Adam, reductions.
// All are coalesced, 16-byte vectorised where alignment allows, warp-shuffle reductions.
#include "kernels.h"

long long g_launch_count = 0;

// ------------------------------------------------------------------------------------------
// embedding gather: one warp per row, 16-byte copies (model.py:111-112)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_embed_gather(const int* __restrict__ ids, long long n,
                                                      const T* __restrict__ table, int D, T* __restrict__ out) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const T* src = table + (long long)ids[row] * D;
    T* dst = out + row * D;
    constexpr int VEC = 16 / sizeof(T);
    if (D % VEC == 0) {
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
        int nv = D / VEC;
        for (int i = lane; i < nv; i += 32) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = lane; i < D; i += 32) dst[i] = src[i];
    }
}
void launch_embed_gather_f32(const int* ids, long long n, const float* table, int D, float* out, cudaStream_t s) {
    if (n <= 0) return;
    k_embed_gather<float><<<cdiv(n, 8), 256, 0, s>>>(ids, n, table, D, out);
    COUNT_LAUNCH();
}
void launch_embed_gather_bf16(const int* ids, long long n, const bf16* table, int D, bf16* out, cudaStream_t s) {
    if (n <= 0) return;
    k_embed_gather<bf16><<<cdiv(n, 8), 256, 0, s>>>(ids, n, table, D, out);
    COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// embedding gradient scatter-add (IndexedSlices part of dE): one warp per row, vector atomics
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_embed_scatter_add(const int* __restrict__ ids, long long n,
                                                           const float* __restrict__ dx, int ld, int D,
                                                           float* __restrict__ g) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* src = dx + row * ld;
    float* dst = g + (long long)ids[row] * D;
    if ((D & 3) == 0 && (ld & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = lane; i < D / 4; i += 32) atomicAdd(d4 + i, s4[i]);
    } else {
        for (int i = lane; i < D; i += 32) atomicAdd(dst + i, src[i]);
    }
}
void launch_embed_scatter_add(const int* ids, long long n, const float* dx, int ld, int D, float* g, cudaStream_t s) {
    if (n <= 0) return;
    k_embed_scatter_add<<<cdiv(n, 8), 256, 0, s>>>(ids, n, dx, ld, D, g);
    COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// row gather / scatter (final-state gather model.py:135, decoder state fan-out model.py:159)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_gather(const float* __restrict__ in_f, const bf16* __restrict__ in_h, int ld_in,
                                                    const int* __restrict__ src_idx, float* __restrict__ out_f,
                                                    bf16* __restrict__ out_h, int ld_out, const int* __restrict__ dst_idx,
                                                    int n, int cols) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    long long si = src_idx ? src_idx[row] : row, di = dst_idx ? dst_idx[row] : row;
    for (int i = lane; i < cols; i += 32) {
        float v = in_f ? in_f[si * ld_in + i] : __bfloat162float(in_h[si * ld_in + i]);
        if (out_f) out_f[di * ld_out + i] = v;
        if (out_h) out_h[di * ld_out + i] = __float2bfloat16(v);
    }
}
void launch_row_gather(const float* in_f, const bf16* in_h, int ld_in, const int* src_idx, float* out_f, bf16* out_h,
                       int ld_out, const int* dst_idx, int n, int cols, cudaStream_t s) {
    if (n <= 0) return;
    k_row_gather<<<cdiv(n, 8), 256, 0, s>>>(in_f, in_h, ld_in, src_idx, out_f, out_h, ld_out, dst_idx, n, cols);
    COUNT_LAUNCH();
}
__global__ void __launch_bounds__(256) k_row_scatter(const float* __restrict__ in, int ld_in, float* __restrict__ out,
                                                     int ld_out, const int* __restrict__ dst_idx, int n, int cols, int acc) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    long long di = dst_idx ? dst_idx[row] : row;
    for (int i = lane; i < cols; i += 32) {
        float v = in[(long long)row * ld_in + i];
        if (acc) out[di * ld_out + i] += v; else out[di * ld_out + i] = v;   // dst rows are unique
    }
}
void launch_row_scatter(const float* in, int ld_in, float* out, int ld_out, const int* dst_idx, int n, int cols,
                        int accumulate, cudaStream_t s) {
    if (n <= 0) return;
    k_row_scatter<<<cdiv(n, 8), 256, 0, s>>>(in, ld_in, out, ld_out, dst_idx, n, cols, accumulate);
    COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// latent: reparameterisation + KL (model.py:147-156,182-184) and its gradient
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t step, uint32_t row, uint32_t col) {
    uint32_t c[4] = {row, col >> 2, (uint32_t)PHILOX_STREAM_EPS, (uint32_t)(seed >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)step);
    int pair = (col >> 1) & 1;
    float u1 = ((float)(c[2 * pair] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(c[2 * pair + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    return (col & 1) ? rad * sn : rad * cs;
}
__global__ void __launch_bounds__(256) k_latent_fwd(const float* __restrict__ mulv, const float* __restrict__ eps_in, int b,
                                                    int R, int train, uint64_t seed, uint64_t step, long long row0,
                                                    const int* __restrict__ row_ids, float* __restrict__ eps_used, float* __restrict__ z_f,
                                                    bf16* __restrict__ z_h, float* __restrict__ kld_samp,
                                                    double* __restrict__ stats) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float k = 0.f;
    if (i < (long long)b * R) {
        int r = (int)(i / R), c = (int)(i % R);
        float mu = mulv[(long long)r * 2 * R + c], lv = mulv[(long long)r * 2 * R + R + c];
        float z = mu;
        if (train) {
            float e = eps_in ? eps_in[i] : philox_normal(seed, step, row_ids ? (uint32_t)row_ids[r] : (uint32_t)(row0 + r), (uint32_t)c);
            eps_used[i] = e;
            z = mu + expf(0.5f * lv) * e;
        }
        if (z_f) z_f[i] = z;
        if (z_h) z_h[i] = __float2bfloat16(z);
        k = 0.5f * (mu * mu + expf(lv) - lv - 1.0f);
        if (kld_samp) kld_samp[i] = k;
    }
    k = warp_sum(k);
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += sm[w];
        atomicAdd(stats + 2, (double)t);
    }
}
void launch_latent_fwd(const float* mulv, const float* eps_in, int b, int R, int train, uint64_t seed, uint64_t step,
                       long long row0, const int* row_ids, float* eps_used, float* z_f, bf16* z_h, float* kld_samp, double* stats,
                       cudaStream_t s) {
    k_latent_fwd<<<cdiv((long long)b * R, 256), 256, 0, s>>>(mulv, eps_in, b, R, train, seed, step, row0, row_ids, eps_used, z_f, z_h,
                                                             kld_samp, stats);
    COUNT_LAUNCH();
}
__global__ void __launch_bounds__(256) k_latent_bwd(const float* __restrict__ dz, const float* __restrict__ mulv,
                                                    const float* __restrict__ eps, int b, int R, int train, float a,
                                                    float* __restrict__ df, bf16* __restrict__ dh) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)b * R) return;
    int r = (int)(i / R), c = (int)(i % R);
    float mu = mulv[(long long)r * 2 * R + c], lv = mulv[(long long)r * 2 * R + R + c];
    float d = dz[i];
    float dmu = d + a * mu;
    float dlv = a * 0.5f * (expf(lv) - 1.0f);
    if (train) dlv += d * 0.5f * expf(0.5f * lv) * eps[i];
    long long o = (long long)r * 2 * R + c;
    if (df) { df[o] = dmu; df[o + R] = dlv; }
    if (dh) { dh[o] = __float2bfloat16(dmu); dh[o + R] = __float2bfloat16(dlv); }
}
void launch_latent_bwd(const float* dz, const float* mulv, const float* eps, int b, int R, int train, float a,
                       float* dmulv_f, bf16* dmulv_h, cudaStream_t s) {
    k_latent_bwd<<<cdiv((long long)b * R, 256), 256, 0, s>>>(dz, mulv, eps, b, R, train, a, dmulv_f, dmulv_h);
    COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// fused softmax cross entropy (model.py:170-181) + gradient, one CTA per row.
// The row is staged once in shared memory (16-byte loads), reduced with warp shuffles, and the
// gradient (softmax - onehot) * gscale is written back in place: 1 read + 1 write of N*V.
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(256) k_ce(T* __restrict__ logits, int ld, const int* __restrict__ labels, int V,
                                            float gscale, int write_grad, float* __restrict__ loss_samp,
                                            float* __restrict__ err_samp, int* __restrict__ pred, double* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* row = reinterpret_cast<T*>(smem_raw);
    __shared__ float red_v[8];
    __shared__ int red_i[8];
    const long long r = blockIdx.x;
    T* g = logits + r * ld;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int VEC = 16 / sizeof(T);
    const bool vec_ok = (V % VEC == 0) && (ld % VEC == 0);
    // pass 0: stage row, track max / argmax (ties -> lowest index, like tf.argmax)
    float mx = -INFINITY;
    int mi = 0x7fffffff;
    if (vec_ok) {
        const int4* g4 = reinterpret_cast<const int4*>(g);
        int4* r4 = reinterpret_cast<int4*>(row);
        for (int i = tid; i < V / VEC; i += 256) {
            int4 v = g4[i];
            r4[i] = v;
            const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float f = to_f<T>(e[j]);
                if (f > mx) { mx = f; mi = i * VEC + j; }
            }
        }
    } else {
        for (int i = tid; i < V; i += 256) {
            T v = g[i];
            row[i] = v;
            float f = to_f<T>(v);
            if (f > mx) { mx = f; mi = i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
    }
    if (lane == 0) { red_v[wid] = mx; red_i[wid] = mi; }
    __syncthreads();
    mx = red_v[0]; mi = red_i[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
        float ov = red_v[w]; int oi = red_i[w];
        if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
    }
    __syncthreads();
    // pass 1: sum exp
    float sum = 0.f;
    for (int i = tid; i < V; i += 256) sum += __expf(to_f<T>(row[i]) - mx);
    sum = warp_sum(sum);
    if (lane == 0) red_v[wid] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red_v[w];
    const int lab = labels ? labels[r] : -1;
    if (tid == 0) {
        if (pred) pred[r] = mi;
        if (labels) {
            float loss = logf(sum) + mx - to_f<T>(row[lab]);
            float err = (mi != lab) ? 1.f : 0.f;
            if (loss_samp) loss_samp[r] = loss;
            if (err_samp) err_samp[r] = err;
            atomicAdd(stats + 0, (double)loss);
            atomicAdd(stats + 1, (double)err);
        }
    }
    if (!write_grad) return;
    const float inv = gscale / sum;
    if (vec_ok) {
        int4* g4 = reinterpret_cast<int4*>(g);
        const int4* r4 = reinterpret_cast<const int4*>(row);
        for (int i = tid; i < V / VEC; i += 256) {
            int4 v = r4[i];
            T* e = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float p = __expf(to_f<T>(e[j]) - mx) * inv;
                if (i * VEC + j == lab) p -= gscale;
                e[j] = from_f<T>(p);
            }
            g4[i] = v;
        }
    } else {
        for (int i = tid; i < V; i += 256) {
            float p = __expf(to_f<T>(row[i]) - mx) * inv;
            if (i == lab) p -= gscale;
            g[i] = from_f<T>(p);
        }
    }
}

// Register-resident fast path for bf16 logits (V = NV * 2048): every thread keeps its NV 16-byte vectors PACKED in
// registers (16 * NV bytes per thread, so 6 rows are resident per SM at V = 8192 and their loads overlap the math of
// the others); a row is read from HBM once and written once, no shared-memory staging pass.  Instruction budget per
// element (the kernel is issue-bound long before it is HBM-bound if written naively): packed bf16x2 max (0.5),
// packed equality for the argmax (0.5), unpack + FFMA + MUFU.EX2 + FADD + pack of exp(x - max) (4.5), unpack + FMUL +
// pack of the gradient (2.5).  exp() is evaluated ONCE per element (MUFU runs 16/clk/SM: twice would cost 1024
// cycles per row against a 1456-cycle HBM budget); it is kept as bf16 between the two passes, so the gradient is
// rounded twice (<= 1 bf16 ulp); the loss uses the fp32 sum.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int NV, int MINB>
__global__ void __launch_bounds__(256, MINB) k_ce_reg(bf16* __restrict__ logits, int ld, const int* __restrict__ labels, float gscale,
                                                      int write_grad, float* __restrict__ loss_samp, float* __restrict__ err_samp,
                                                      int* __restrict__ pred, double* __restrict__ stats) {
    __shared__ float red_v[8];
    __shared__ int red_i[8];
    const long long r = blockIdx.x;
    uint4* g4 = reinterpret_cast<uint4*>(logits + r * ld);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t w[NV][4];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const uint4 v = g4[j * 256 + tid];
        w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
    }
    const int lab = labels ? labels[r] : -1;
    // the thread that holds the label element finishes the row (loss, error flag, label gradient in fp32)
    const bool owner = lab >= 0 ? ((lab >> 3) & 255) == tid : tid == 0;
    float x_lab = 0.f;
    if (owner && lab >= 0) x_lab = __bfloat162float(logits[r * ld + lab]);   // read before the row is overwritten
    // ---- max (packed bf16x2)
    __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&w[0][0]);
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) m2 = __hmax2(m2, *reinterpret_cast<const __nv_bfloat162*>(&w[j][q]));
    const float mt = fmaxf(__low2float(m2), __high2float(m2));
    float mx = mt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red_v[wid] = mx;
    __syncthreads();
    mx = red_v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red_v[i]);
    __syncthreads();
    // ---- argmax (ties -> lowest index, like tf.argmax): only the threads whose own maximum is the row maximum look
    int mi = 0x7fffffff;
    if (mt == mx) {
        const __nv_bfloat162 mx2 = __float2bfloat162_rn(mx);   // exact: mx is one of the bf16 inputs
#pragma unroll
        for (int j = NV - 1; j >= 0; --j)
#pragma unroll
            for (int q = 3; q >= 0; --q) {   // descending, so the last hit is the lowest index
                const unsigned eq = __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w[j][q]), mx2);
                if (eq) mi = (j * 256 + tid) * 8 + 2 * q + ((eq & 0xffffu) ? 0 : 1);
            }
    }
    // ---- exp(x - max) once per element, fp32 sum; kept packed as bf16 for the gradient pass
    constexpr float L2E = 1.4426950408889634f;
    const float nmx = -mx * L2E;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(w[j][q] << 16), L2E, nmx));
            const float e1 = ex2_approx(fmaf(__uint_as_float(w[j][q] & 0xffff0000u), L2E, nmx));
            sum0 += e0;
            sum1 += e1;
            const __nv_bfloat162 pk = __floats2bfloat162_rn(e0, e1);
            w[j][q] = *reinterpret_cast<const uint32_t*>(&pk);
        }
    float sum = warp_sum(sum0 + sum1);
    mi = __reduce_min_sync(0xffffffffu, mi);
    if (lane == 0) { red_v[wid] = sum; red_i[wid] = mi; }
    __syncthreads();
    sum = 0.f;
    mi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sum += red_v[i]; mi = min(mi, red_i[i]); }
    const float inv = gscale / sum;
    if (write_grad) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(w[j][q] << 16) * inv,
                                                                 __uint_as_float(w[j][q] & 0xffff0000u) * inv);
                w[j][q] = *reinterpret_cast<const uint32_t*>(&pk);
            }
            g4[j * 256 + tid] = make_uint4(w[j][0], w[j][1], w[j][2], w[j][3]);
        }
    }
    if (owner) {
        if (pred) pred[r] = mi;
        if (lab >= 0) {
            const float loss = logf(sum) + mx - x_lab;
            const float err = (mi != lab) ? 1.f : 0.f;
            if (loss_samp) loss_samp[r] = loss;
            if (err_samp) err_samp[r] = err;
            atomicAdd(stats + 0, (double)loss);
            atomicAdd(stats + 1, (double)err);
            // label element: softmax - 1 from the fp32 exp (same thread stored the vector above: program order)
            if (write_grad) logits[r * ld + lab] = __float2bfloat16(ex2_approx(fmaf(x_lab, L2E, nmx)) * inv - gscale);
        }
    }
}
template <int NV, int MINB>
static void launch_ce_reg(bf16* logits, int ld, const int* labels, long long n, float gscale, int write_grad, float* loss_samp,
                          float* err_samp, int* pred, double* stats, cudaStream_t s) {
    k_ce_reg<NV, MINB><<<(unsigned)n, 256, 0, s>>>(logits, ld, labels, gscale, write_grad, loss_samp, err_samp, pred, stats);
    COUNT_LAUNCH();
}
template <typename T>
static void launch_ce(T* logits, int ld, const int* labels, long long n, int V, float gscale, int write_grad,
                      float* loss_samp, float* err_samp, int* pred, double* stats, cudaStream_t s) {
    if (n <= 0) return;
    size_t smem = (size_t)V * sizeof(T);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(k_ce<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    k_ce<T><<<(unsigned)n, 256, smem, s>>>(logits, ld, labels, V, gscale, write_grad, loss_samp, err_samp, pred, stats);
    COUNT_LAUNCH();
}
void launch_ce_f32(float* logits, int ld, const int* labels, long long n, int V, float gscale, int write_grad,
                   float* loss_samp, float* err_samp, int* pred, double* stats, cudaStream_t s) {
    launch_ce<float>(logits, ld, labels, n, V, gscale, write_grad, loss_samp, err_samp, pred, stats, s);
}
void launch_ce_bf16(bf16* logits, int ld, const int* labels, long long n, int V, float gscale, int write_grad,
                    float* loss_samp, float* err_samp, int* pred, double* stats, cudaStream_t s) {
    if (n > 0 && V % 2048 == 0 && ld % 8 == 0 && ((uintptr_t)logits & 15) == 0 && !getenv("ARGSIM_CE_SMEM")) {
        switch (V / 2048) {
            case 1: return launch_ce_reg<1, 8>(logits, ld, labels, n, gscale, write_grad, loss_samp, err_samp, pred, stats, s);
            case 2: return launch_ce_reg<2, 6>(logits, ld, labels, n, gscale, write_grad, loss_samp, err_samp, pred, stats, s);
            case 4: return launch_ce_reg<4, 5>(logits, ld, labels, n, gscale, write_grad, loss_samp, err_samp, pred, stats, s);
            case 8: return launch_ce_reg<8, 3>(logits, ld, labels, n, gscale, write_grad, loss_samp, err_samp, pred, stats, s);
            case 16: return launch_ce_reg<16, 2>(logits, ld, labels, n, gscale, write_grad, loss_samp, err_samp, pred, stats, s);
            default: break;
        }
    }
    launch_ce<bf16>(logits, ld, labels, n, V, gscale, write_grad, loss_samp, err_samp, pred, stats, s);
}

// ------------------------------------------------------------------------------------------
// Adam, TF-1 form (model.py:189): theta -= lr_t * m / (sqrt(v) + eps).  28 B/param (+2 shadow).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, bf16* __restrict__ shadow, long long n4, long long n,
                                              float lr_t, float b1, float b2, float eps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* Vv = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            M[j] = b1 * M[j] + (1.f - b1) * G[j];
            Vv[j] = b2 * Vv[j] + (1.f - b2) * G[j] * G[j];
            P[j] -= lr_t * M[j] / (sqrtf(Vv[j]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (shadow) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(shadow)[i] = pk;
        }
    }
    // tail (n not multiple of 4)
    long long t = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        float mm = b1 * m[t] + (1.f - b1) * g[t];
        float vv = b2 * v[t] + (1.f - b2) * g[t] * g[t];
        float pp = p[t] - lr_t * mm / (sqrtf(vv) + eps);
        m[t] = mm; v[t] = vv; p[t] = pp;
        if (shadow) shadow[t] = __float2bfloat16(pp);
    }
}
void launch_adam(float* p, const float* g, float* m, float* v, bf16* shadow, long long n, float lr_t, float b1,
                 float b2, float eps, cudaStream_t s) {
    if (n <= 0) return;
    long long n4 = n / 4;
    int blocks = (int)std::min<long long>(cdiv(std::max<long long>(n4, 1), 256), 148LL * 16);
    k_adam<<<blocks, 256, 0, s>>>(p, g, m, v, shadow, n4, n, lr_t, b1, b2, eps);
    COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_colsum(const T* __restrict__ in, int ld, long long rows, int cols,
                                                float* __restrict__ out) {
    __shared__ float sm[8][33];
    int c = blockIdx.x * 32 + (threadIdx.x & 31);
    int ry = threadIdx.x >> 5;
    long long chunk = (rows + gridDim.y - 1) / gridDim.y;
    long long r0 = (long long)blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
    float acc = 0.f;
    if (c < cols)
        for (long long r = r0 + ry; r < r1; r += 8) acc += to_f<T>(in[r * ld + c]);
    sm[ry][threadIdx.x & 31] = acc;
    __syncthreads();
    if (ry == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += sm[j][threadIdx.x];
        atomicAdd(out + c, t);
    }
}
template <typename T>
static void launch_colsum(const T* in, int ld, long long rows, int cols, float* out, cudaStream_t s, int accumulate) {
    if (!accumulate) CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * cols, s));
    if (rows <= 0) return;
    int gy = (int)std::max<long long>(1, std::min<long long>(64, rows / 64));
    dim3 grid(cdiv(cols, 32), gy);
    k_colsum<T><<<grid, 256, 0, s>>>(in, ld, rows, cols, out);
    COUNT_LAUNCH();
}
void launch_colsum_f32(const float* in, int ld, long long rows, int cols, float* out, cudaStream_t s, int accumulate) {
    launch_colsum<float>(in, ld, rows, cols, out, s, accumulate);
}
void launch_colsum_bf16(const bf16* in, int ld, long long rows, int cols, float* out, cudaStream_t s, int accumulate) {
    launch_colsum<bf16>(in, ld, rows, cols, out, s, accumulate);
}

__global__ void __launch_bounds__(256) k_cast_bf16(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __float2bfloat16(in[i]);
}
void launch_cast_bf16(const float* in, bf16* out, long long n, cudaStream_t s) {
    if (n <= 0) return;
    int blocks = (int)std::min<long long>(cdiv(n, 256), 148LL * 16);
    k_cast_bf16<<<blocks, 256, 0, s>>>(in, out, n);
    COUNT_LAUNCH();
}
__global__ void __launch_bounds__(256) k_cast_f32(const bf16* __restrict__ in, float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}
void launch_cast_f32(const bf16* in, float* out, long long n, cudaStream_t s) {
    if (n <= 0) return;
    int blocks = (int)std::min<long long>(cdiv(n, 256), 148LL * 16);
    k_cast_f32<<<blocks, 256, 0, s>>>(in, out, n);
    COUNT_LAUNCH();
}
// greedy decode (model.py:204-219), device-resident loop: one step's bookkeeping.  Appends pred to the output,
// feeds it back as the next lead, and raises the stop flag when every row emitted eos (the reference breaks BEFORE
// appending that step).  flag[0] = stopped, flag[1] = number of steps kept.
__global__ void __launch_bounds__(256) k_decode_advance(const int* __restrict__ pred, int b, int eos, int* __restrict__ out,
                                                        int* __restrict__ lead, int* __restrict__ flag) {
    // flag[0] = stopped, flag[1] = steps kept so far (= index of this step), flag[2] = step budget: nothing here depends on
    // the host, so the launch can be replayed from a CUDA graph
    const int t = flag[1];
    if (flag[0] || t >= flag[2]) return;
    int all = 1;
    for (int i = threadIdx.x; i < b; i += 256) all &= (pred[i] == eos);
    all = __syncthreads_and(all);
    if (all) {
        if (threadIdx.x == 0) flag[0] = 1;
        return;
    }
    for (int i = threadIdx.x; i < b; i += 256) {
        const int v = pred[i];
        out[(size_t)t * b + i] = v;
        lead[i] = v;
    }
    if (threadIdx.x == 0) flag[1] = t + 1;
}
void launch_decode_advance(const int* pred, int b, int eos, int* out, int* lead, int* flag, cudaStream_t s) {
    k_decode_advance<<<1, 256, 0, s>>>(pred, b, eos, out, lead, flag);
    COUNT_LAUNCH();
}
__global__ void __launch_bounds__(256) k_fill_i32(int* p, long long n, int v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill_i32(int* p, long long n, int v, cudaStream_t s) {
    if (n <= 0) return;
    k_fill_i32<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(p, n, v);
    COUNT_LAUNCH();
}
__global__ void __launch_bounds__(256) k_fill(float* p, long long n, float v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
void launch_fill(float* p, long long n, float v, cudaStream_t s) {
    if (n <= 0) return;
    int blocks = (int)std::min<long long>(cdiv(n, 256), 148LL * 16);
    k_fill<<<blocks, 256, 0, s>>>(p, n, v);
    COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// attentive=true (src/model.py:18-45,136-145 with the reshape at :36 and the feature axis of the affines repaired):
// one query per sequence (its final state) attends over that sequence's own top-layer outputs, 8 heads, then
// h <- layer_norm(h + p(attend)).  K and V live in the packed time-major row layout of the encoder (row of step t of
// sorted sequence j = off[t] + j), so the mask of the reference (log of 0/1 over padding) is the loop bound here.
// ------------------------------------------------------------------------------------------
namespace {
constexpr int ATT_NT = 128;
__device__ __forceinline__ float block_sum_att(float v, float* red) {   // ATT_NT threads; red: 4 floats + 1
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}
__device__ __forceinline__ float block_max_att(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}
// the sequence whose last step sits in packed row `last`: number of steps and sorted position
__device__ __forceinline__ void seq_of_last_row(const int* __restrict__ off, int Tmax, int last, int* len, int* j) {
    int lo = 0, hi = Tmax - 1;   // largest t with off[t] <= last
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= last) lo = mid; else hi = mid - 1;
    }
    *len = lo + 1;
    *j = last - off[lo];
}
}  // namespace

// grid (b, heads): scores over the sequence's own steps, softmax, weighted sum of V.  prob: (S, heads) packed.
__global__ void __launch_bounds__(ATT_NT) k_attn_fwd(const float* __restrict__ q, const float* __restrict__ Kf, const float* __restrict__ Vf,
                                                     int dim, int heads, const int* __restrict__ off, int Tmax,
                                                     const int* __restrict__ last_row, float* __restrict__ prob,
                                                     float* __restrict__ y_f, bf16* __restrict__ y_h) {
    extern __shared__ float att_sm[];
    const int c = dim / heads;
    float* qs = att_sm;            // c
    float* sc = att_sm + c;        // Tmax
    float* red = sc + Tmax;        // 4
    const int i = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    int len, j;
    seq_of_last_row(off, Tmax, last_row[i], &len, &j);
    for (int k = tid; k < c; k += ATT_NT) qs[k] = q[(size_t)i * dim + hd * c + k];
    __syncthreads();
    const float scale = rsqrtf((float)c);
    for (int t = wrp; t < len; t += ATT_NT / 32) {
        const float* kr = Kf + (size_t)(off[t] + j) * dim + hd * c;
        float d = 0.f;
        for (int k = lane; k < c; k += 32) d += qs[k] * kr[k];
        d = warp_sum(d);
        if (lane == 0) sc[t] = d * scale;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int t = tid; t < len; t += ATT_NT) mx = fmaxf(mx, sc[t]);
    mx = block_max_att(mx, red);
    float sm = 0.f;
    for (int t = tid; t < len; t += ATT_NT) { const float e = __expf(sc[t] - mx); sc[t] = e; sm += e; }
    sm = block_sum_att(sm, red);
    const float inv = 1.0f / sm;
    for (int t = tid; t < len; t += ATT_NT) {
        const float a = sc[t] * inv;
        sc[t] = a;
        prob[(size_t)(off[t] + j) * heads + hd] = a;
    }
    __syncthreads();
    for (int k = tid; k < c; k += ATT_NT) {
        float acc = 0.f;
        for (int t = 0; t < len; ++t) acc += sc[t] * Vf[(size_t)(off[t] + j) * dim + hd * c + k];
        const size_t o = (size_t)i * dim + hd * c + k;
        y_f[o] = acc;
        if (y_h) y_h[o] = __float2bfloat16(acc);
    }
}
void launch_attn_fwd(const float* q, const float* K, const float* V, int b, int dim, int heads, const int* off, int Tmax,
                     const int* last_row, float* prob, float* y_f, bf16* y_h, cudaStream_t s) {
    if (b <= 0) return;
    const size_t smem = sizeof(float) * (dim / heads + Tmax + 8);
    static bool configured = false;
    if (!configured) {   // sequences of more than ~12,000 steps need more than the default 48 KB
        CUDA_CHECK(cudaFuncSetAttribute(k_attn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    if (smem > 200 * 1024) throw std::runtime_error("attention: sequence too long for the per-sequence score buffer");
    k_attn_fwd<<<dim3(b, heads), ATT_NT, smem, s>>>(q, K, V, dim, heads, off, Tmax, last_row, prob, y_f, y_h);
    COUNT_LAUNCH();
}

// grid (b, heads): every (packed row, head) of dK / dV belongs to exactly one block -- no atomics, no memset
__global__ void __launch_bounds__(ATT_NT) k_attn_bwd(const float* __restrict__ dy, const float* __restrict__ q, const float* __restrict__ Kf,
                                                     const float* __restrict__ Vf, const float* __restrict__ prob, int dim, int heads,
                                                     const int* __restrict__ off, int Tmax, const int* __restrict__ last_row,
                                                     float* __restrict__ dq_f, bf16* __restrict__ dq_h, float* __restrict__ dK_f,
                                                     bf16* __restrict__ dK_h, float* __restrict__ dV_f, bf16* __restrict__ dV_h) {
    extern __shared__ float att_sm[];
    const int c = dim / heads;
    float* qs = att_sm;            // c
    float* dys = qs + c;           // c
    float* ds = dys + c;           // Tmax: d score
    float* pa = ds + Tmax;         // Tmax: a
    float* red = pa + Tmax;        // 4
    const int i = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    int len, j;
    seq_of_last_row(off, Tmax, last_row[i], &len, &j);
    for (int k = tid; k < c; k += ATT_NT) {
        qs[k] = q[(size_t)i * dim + hd * c + k];
        dys[k] = dy[(size_t)i * dim + hd * c + k];
    }
    __syncthreads();
    const float scale = rsqrtf((float)c);
    for (int t = wrp; t < len; t += ATT_NT / 32) {   // da[t] = dy . V[t]
        const float* vr = Vf + (size_t)(off[t] + j) * dim + hd * c;
        float d = 0.f;
        for (int k = lane; k < c; k += 32) d += dys[k] * vr[k];
        d = warp_sum(d);
        if (lane == 0) { ds[t] = d; pa[t] = prob[(size_t)(off[t] + j) * heads + hd]; }
    }
    __syncthreads();
    float dot = 0.f;
    for (int t = tid; t < len; t += ATT_NT) dot += pa[t] * ds[t];
    dot = block_sum_att(dot, red);
    for (int t = tid; t < len; t += ATT_NT) ds[t] = pa[t] * (ds[t] - dot) * scale;   // d (q.k) including the c^-1/2
    __syncthreads();
    for (int k = tid; k < c; k += ATT_NT) {
        float acc = 0.f;
        const float qk = qs[k], dyk = dys[k];
        for (int t = 0; t < len; ++t) {
            const size_t o = (size_t)(off[t] + j) * dim + hd * c + k;
            acc += ds[t] * Kf[o];
            const float dk = ds[t] * qk, dv = pa[t] * dyk;
            dK_f[o] = dk; dV_f[o] = dv;
            if (dK_h) { dK_h[o] = __float2bfloat16(dk); dV_h[o] = __float2bfloat16(dv); }
        }
        const size_t o = (size_t)i * dim + hd * c + k;
        dq_f[o] = acc;
        if (dq_h) dq_h[o] = __float2bfloat16(acc);
    }
}
void launch_attn_bwd(const float* dy, const float* q, const float* K, const float* V, const float* prob, int b, int dim, int heads,
                     const int* off, int Tmax, const int* last_row, float* dq_f, bf16* dq_h, float* dK_f, bf16* dK_h, float* dV_f,
                     bf16* dV_h, cudaStream_t s) {
    if (b <= 0) return;
    const size_t smem = sizeof(float) * (2 * (dim / heads) + 2 * Tmax + 8);
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(k_attn_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    if (smem > 200 * 1024) throw std::runtime_error("attention: sequence too long for the per-sequence score buffers");
    k_attn_bwd<<<dim3(b, heads), ATT_NT, smem, s>>>(dy, q, K, V, prob, dim, heads, off, Tmax, last_row, dq_f, dq_h, dK_f, dK_h, dV_f, dV_h);
    COUNT_LAUNCH();
}

// tf.contrib.layers.layer_norm (src/model.py:11,139): over the last axis, biased variance, epsilon 1e-12, gamma / beta.
// x = h + p (the residual is formed here).  Saves xhat and 1/sigma for the backward.
__global__ void __launch_bounds__(ATT_NT) k_resid_ln_fwd(const float* __restrict__ h, const float* __restrict__ pp, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int dim, float* __restrict__ xhat,
                                                         float* __restrict__ rstd, float* __restrict__ out_f, bf16* __restrict__ out_h) {
    __shared__ float red[8];
    const int i = blockIdx.x, tid = threadIdx.x;
    const float* hr = h + (size_t)i * dim; const float* pr = pp + (size_t)i * dim;
    float sm = 0.f;
    for (int k = tid; k < dim; k += ATT_NT) sm += hr[k] + pr[k];
    const float mean = block_sum_att(sm, red) / (float)dim;
    float sq = 0.f;
    for (int k = tid; k < dim; k += ATT_NT) { const float d = hr[k] + pr[k] - mean; sq += d * d; }
    const float var = block_sum_att(sq, red) / (float)dim;
    const float rs = rsqrtf(var + 1e-12f);
    if (tid == 0) rstd[i] = rs;
    for (int k = tid; k < dim; k += ATT_NT) {
        const float xh = (hr[k] + pr[k] - mean) * rs;
        const float o = xh * gamma[k] + beta[k];
        xhat[(size_t)i * dim + k] = xh;
        out_f[(size_t)i * dim + k] = o;
        if (out_h) out_h[(size_t)i * dim + k] = __float2bfloat16(o);
    }
}
void launch_resid_ln_fwd(const float* h, const float* pp, const float* gamma, const float* beta, int b, int dim, float* xhat, float* rstd,
                         float* out_f, bf16* out_h, cudaStream_t s) {
    if (b <= 0) return;
    k_resid_ln_fwd<<<b, ATT_NT, 0, s>>>(h, pp, gamma, beta, dim, xhat, rstd, out_f, out_h);
    COUNT_LAUNCH();
}
// dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dout gamma; also writes dout * xhat (its column sum is d gamma)
__global__ void __launch_bounds__(ATT_NT) k_ln_bwd(const float* __restrict__ dout, const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                   const float* __restrict__ gamma, int dim, float* __restrict__ dx_f, bf16* __restrict__ dx_h,
                                                   float* __restrict__ dgam_rows) {
    __shared__ float red[8];
    const int i = blockIdx.x, tid = threadIdx.x;
    const float* dr = dout + (size_t)i * dim; const float* xr = xhat + (size_t)i * dim;
    float s1 = 0.f, s2 = 0.f;
    for (int k = tid; k < dim; k += ATT_NT) { const float gk = dr[k] * gamma[k]; s1 += gk; s2 += gk * xr[k]; }
    const float m1 = block_sum_att(s1, red) / (float)dim;
    const float m2 = block_sum_att(s2, red) / (float)dim;
    const float rs = rstd[i];
    for (int k = tid; k < dim; k += ATT_NT) {
        const float gk = dr[k] * gamma[k];
        const float dx = rs * (gk - m1 - xr[k] * m2);
        dx_f[(size_t)i * dim + k] = dx;
        if (dx_h) dx_h[(size_t)i * dim + k] = __float2bfloat16(dx);
        dgam_rows[(size_t)i * dim + k] = dr[k] * xr[k];
    }
}
void launch_ln_bwd(const float* dout, const float* xhat, const float* rstd, const float* gamma, int b, int dim, float* dx_f, bf16* dx_h,
                   float* dgam_rows, cudaStream_t s) {
    if (b <= 0) return;
    k_ln_bwd<<<b, ATT_NT, 0, s>>>(dout, xhat, rstd, gamma, dim, dx_f, dx_h, dgam_rows);
    COUNT_LAUNCH();
}
