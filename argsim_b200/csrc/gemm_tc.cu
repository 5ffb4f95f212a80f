// gemm_tc.cu -- bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (sm_100a only):
//   C(M,N) = alpha * A(M,K) . B(N,K)^T (+ bias[N]) (+ C)
// TMA (cp.async.bulk.tensor, 128-byte swizzle) stages 128x64 / 64x128 operand tiles through a
// 4-deep shared-memory ring; ONE thread issues tcgen05.mma (M=128, N=128, K=16) into a TMEM
// accumulator; four epilogue warps read it back with tcgen05.ld and write fp32 and/or bf16.
// Both operands may be K-major (rows x K, K contiguous) or MN-major (K x rows, rows contiguous:
// the wgrad / dgrad shapes), selected by the UMMA descriptors -- no transposes in HBM.
// Split-K (fp32 red.global.add) keeps all 148 SMs busy on the weight-gradient shapes.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane group = warp_id % 4).
//
// Two kernels share the operand pipeline:
//   k_gemm_tc2 (default): PERSISTENT, one CTA per SM walking (m-tile, n-tile, k-split) work units; 128 x 256 tiles
//     (96 B/clk of shared-memory operand reads per MMA instead of the 128 B/clk a 128 x 128 tile needs -- the port
//     limit); the fp32 accumulator is DOUBLE BUFFERED in TMEM (2 x 256 columns) so the epilogue of unit i runs under
//     the MMAs of unit i+1; the epilogue goes TMEM -> registers -> 128B-swizzled shared memory -> TMA store
//     (cp.async.bulk.tensor, or cp.reduce.async.bulk.tensor .add for accumulate / split-K), so global writes are
//     full 128-byte lines issued by the copy engine instead of 16-byte-per-row scattered stores.
//   k_gemm_tc (v1): one 128 x 128 tile per CTA, direct stores; kept for outputs that want fp32 AND bf16 copies or are
//     not 16-byte aligned, and as the A/B baseline (ARGSIM_GEMM_V1=1).
#include "kernels.h"
#include <cuda.h>
#include <mutex>

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 128;
constexpr int NTHREADS = 192;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_ready = false, g_tried = false;
int g_num_sms = 148;

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a pipeline bug must not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// one lane of a CONVERGED warp.  Guarding tcgen05 / TMA issue with `lane == 0` makes ptxas wrap every such instruction in
// an ELECT / BRA.U.ANY loop (operands not provably warp-uniform, ~59 cycles per instruction measured in gru_tc.cu);
// a predicate that comes from elect.sync lets it issue them back to back from uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
// K-major tile  (rows x 64 bf16, 128 B per row)  : SBO = 1024 (8-row group), LBO unused (=1)
// MN-major tile (64 k-rows x 64 mn, 128 B per k) : SBO = 1024 (8-k group),  LBO = BK*128 (next 64 mn)
template <int MN_MAJOR>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint64_t lbo = MN_MAJOR ? (uint64_t)((BK * 128) >> 4) : 1ull;
    const uint64_t sbo = 1024 >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
// a_major bit 15, b_major bit 16, N>>3 at [17,23), M>>4 at [24,29)
template <int A_MN, int B_MN>
__device__ __forceinline__ uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) | ((uint32_t)(BN >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

// epilogue modes
enum { EPI_STORE = 0, EPI_ACCUM = 1, EPI_ATOMIC = 2 };

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ Cf,
          bf16* __restrict__ Ch, int ldc, int M, int N, int K, float alpha, const float* __restrict__ bias, int mode,
          int kb_per_split, int vec_ok) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    // barriers: full[STAGES], empty[STAGES], tmem_full ; then the TMEM base-address slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb_total = (K + BK - 1) / BK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int nkb = min(nkb_total, kb0 + kb_per_split) - kb0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_expect_tx(full_bar(s), STAGE_BYTES);
                const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
                const int k = (kb0 + i) * BK;
                if (!A_MN) {
                    tma_load_2d(sa, &tmA, k, m0, full_bar(s));
                } else {
                    tma_load_2d(sa, &tmA, m0, k, full_bar(s));
                    tma_load_2d(sa + BK * 128, &tmA, m0 + 64, k, full_bar(s));
                }
                if (!B_MN) {
                    tma_load_2d(sb, &tmB, k, n0, full_bar(s));
                } else {
                    tma_load_2d(sb, &tmB, n0, k, full_bar(s));
                    tma_load_2d(sb + BK * 128, &tmB, n0 + 64, k, full_bar(s));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t idesc = make_idesc<A_MN, B_MN>();
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
                for (int k16 = 0; k16 < BK / 16; ++k16) {
                    const uint64_t da = make_desc<A_MN>(sa + (A_MN ? k16 * 2048 : k16 * 32));
                    const uint64_t db = make_desc<B_MN>(sb + (B_MN ? k16 * 2048 : k16 * 32));
                    tc_mma(tmem_base, da, db, idesc, (i > 0 || k16 > 0) ? 1u : 0u);
                }
                tc_commit(empty_bar(s));  // smem slot reusable once these MMAs retire
            }
            tc_commit(tmem_full_bar);
        }
    } else if (nkb > 0) {
        // ---- epilogue: TMEM -> registers -> global.  thread <-> one accumulator row (TMEM lane)
        const int lg = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int row = m0 + lg * 32 + lane;
        const bool add_bias = bias != nullptr && (mode != EPI_ATOMIC || blockIdx.z == 0);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c * 32), r);
            const int col0 = n0 + c * 32;
            if (row >= M || col0 >= N) continue;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                v[i] = alpha * __uint_as_float(r[i]);
                if (add_bias && col0 + i < N) v[i] += __ldg(bias + col0 + i);
            }
            const long long o = (long long)row * ldc + col0;
            if (mode == EPI_ATOMIC) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (col0 + i < N) atomicAdd(Cf + o + i, v[i]);
                continue;
            }
            const bool full = vec_ok && (col0 + 32 <= N);
            if (Cf) {
                if (full) {
                    float4* dst = reinterpret_cast<float4*>(Cf + o);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 x = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        if (mode == EPI_ACCUM) {
                            const float4 y = dst[i];
                            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
                            v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
                        }
                        dst[i] = x;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < N) {
                            if (mode == EPI_ACCUM) v[i] += Cf[o + i];
                            Cf[o + i] = v[i];
                        }
                }
            }
            if (Ch) {
                if (full) {
                    uint4* dst = reinterpret_cast<uint4*>(Ch + o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
                        __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
                        __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                        uint4 x;
                        x.x = *reinterpret_cast<uint32_t*>(&p0); x.y = *reinterpret_cast<uint32_t*>(&p1);
                        x.z = *reinterpret_cast<uint32_t*>(&p2); x.w = *reinterpret_cast<uint32_t*>(&p3);
                        dst[i] = x;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < N) Ch[o + i] = __float2bfloat16(v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// ============================================================================================
// v2: persistent, 128 x 256 tiles, double-buffered TMEM accumulator, TMA-store epilogue
// ============================================================================================
namespace v2 {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SLAB_BYTES = 32 * 128;                 // 32 rows x 128 B: one epilogue warp's TMA-store box
constexpr int EPI_BYTES = 4 * 2 * SLAB_BYTES;        // 4 epilogue warps x 2 slabs
constexpr int BIAS_BYTES = BN * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BIAS_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr int TMEM_COLS = 512;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1, int reduce) {
    if (reduce)
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&p);
}
// MN-major tiles are stored as [mn-block of 64][64 k-rows][64 mn] so LBO = 64 k-rows * 128 B
template <int MN_MAJOR>
__device__ __forceinline__ uint64_t make_desc2(uint32_t saddr) {
    const uint64_t lbo = MN_MAJOR ? (uint64_t)((BK * 128) >> 4) : 1ull;
    const uint64_t sbo = 1024 >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}
template <int A_MN, int B_MN>
__device__ __forceinline__ uint32_t make_idesc2() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) | ((uint32_t)(BN >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

struct Work {
    int tiles_m, tiles_n, splits, kb_per_split, nkb_total;
    int b_first = 0;   // dependent launch with a B operand (weights) that does not depend on the predecessor: its first stages
                       // are loaded before griddepcontrol.wait
    __device__ __forceinline__ int count() const { return tiles_m * tiles_n * splits; }
    __device__ __forceinline__ void decode(int u, int& m0, int& n0, int& kb0, int& nkb, int& split) const {
        const int mt = u % tiles_m, rest = u / tiles_m;   // m fastest: CTAs running together share the B (weight) tile in L2
        const int nt = rest % tiles_n;
        split = rest / tiles_n;
        m0 = mt * BM; n0 = nt * BN;
        kb0 = split * kb_per_split;
        nkb = min(nkb_total, kb0 + kb_per_split) - kb0;
    }
};

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           const __grid_constant__ CUtensorMap tmC, int M, int N, float alpha, const float* __restrict__ bias, int out_bf16,
           int reduce, const Work W) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = base + STAGES * STAGE_BYTES;
    const uint32_t bias_base = epi_base + EPI_BYTES;
    const uint32_t bar_base = bias_base + BIAS_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nunits = W.count();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);   // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    // programmatic dependent launch (no-ops on a plain launch): the kernel behind this one may become resident now, and this
    // one -- when it was launched with the attribute (gemm_tc_pdl) -- touches global memory only after its predecessor is complete
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp != 0) asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer (whole warp walks the loop, one lane issues)
        {
            const bool leader = elect_one();
            auto load_b = [&](int s, int k, int n0) {
                const uint32_t sb = base + s * STAGE_BYTES + A_BYTES;
                if (!B_MN) {
                    tma_load_2d(sb, &tmB, k, n0, full_bar(s));
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, n0 + 64 * j, k, full_bar(s));
                }
            };
            int pre = 0;   // stages whose B tile is already on its way
            if (W.b_first && (int)blockIdx.x < nunits) {
                int m0, n0, kb0, nkb, split;
                W.decode(blockIdx.x, m0, n0, kb0, nkb, split);
                pre = min(nkb, STAGES);
                if (leader)
                    for (int i = 0; i < pre; ++i) {
                        mbar_expect_tx(full_bar(i), STAGE_BYTES);
                        load_b(i, (kb0 + i) * BK, n0);
                    }
                __syncwarp();
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            uint32_t it = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
                int m0, n0, kb0, nkb, split;
                W.decode(u, m0, n0, kb0, nkb, split);
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    const bool b_done = (int)it < pre;
                    if (!b_done) mbar_wait(empty_bar(s), ph ^ 1u);
                    if (leader) {
                        if (!b_done) mbar_expect_tx(full_bar(s), STAGE_BYTES);
                        const uint32_t sa = base + s * STAGE_BYTES;
                        const int k = (kb0 + i) * BK;
                        if (!A_MN) {
                            tma_load_2d(sa, &tmA, k, m0, full_bar(s));
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, m0 + 64 * j, k, full_bar(s));
                        }
                        if (!b_done) load_b(s, k, n0);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer (whole warp waits, one lane issues)
        {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc2<A_MN, B_MN>();
            uint32_t it = 0, tl = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++tl) {
                int m0, n0, kb0, nkb, split;
                W.decode(u, m0, n0, kb0, nkb, split);
                const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
                mbar_wait(tempty_bar(acc), aph ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * BN;
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
                        for (int k16 = 0; k16 < BK / 16; ++k16) {
                            const uint64_t da = make_desc2<A_MN>(sa + (A_MN ? k16 * 2048 : k16 * 32));
                            const uint64_t db = make_desc2<B_MN>(sb + (B_MN ? k16 * 2048 : k16 * 32));
                            tc_mma(tacc, da, db, idesc, (i > 0 || k16 > 0) ? 1u : 0u);
                        }
                        tc_commit(empty_bar(s));
                    }
                    __syncwarp();
                }
                if (leader) tc_commit(tfull_bar(acc));
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (4 warps)
        const int lg = warp & 3;
        const bool leader = elect_one();   // the lane that owns this warp's TMA-store bulk groups
        const int et = threadIdx.x - 64;
        const uint32_t slab0 = epi_base + (uint32_t)(warp - 2) * 2 * SLAB_BYTES;
        const uint32_t row_off = (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        uint32_t slab_sel = 0, tl = 0;
        for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++tl) {
            int m0, n0, kb0, nkb, split;
            W.decode(u, m0, n0, kb0, nkb, split);
            const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
            if (bias) {   // the tile's 256 bias values -> smem (split 0 only adds them)
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int cidx = n0 + et + 128 * j;
                    const float bv = (split == 0 && cidx < N) ? __ldg(bias + cidx) : 0.f;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_base + 4u * (et + 128 * j)), "f"(bv) : "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const uint32_t tacc = tmem_base + acc * BN + ((uint32_t)(lg * 32) << 16);
            const int row0 = m0 + lg * 32;
            const bool rows_live = row0 < M;
#pragma unroll 1
            for (int c = 0; c < BN / 64; ++c) {
                uint32_t r[64];
                tc_ld32_nowait(tacc + c * 64, r);
                tc_ld32_nowait(tacc + c * 64 + 32, r + 32);
                tc_wait_ld();
                const int col0 = n0 + c * 64;
                if (!rows_live || col0 >= N) continue;   // warp-uniform
                if (bias) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        float4 b4;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(bias_base + 4u * (c * 64 + 4 * q)));
                        r[4 * q] = __float_as_uint(fmaf(alpha, __uint_as_float(r[4 * q]), b4.x));
                        r[4 * q + 1] = __float_as_uint(fmaf(alpha, __uint_as_float(r[4 * q + 1]), b4.y));
                        r[4 * q + 2] = __float_as_uint(fmaf(alpha, __uint_as_float(r[4 * q + 2]), b4.z));
                        r[4 * q + 3] = __float_as_uint(fmaf(alpha, __uint_as_float(r[4 * q + 3]), b4.w));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) r[i] = __float_as_uint(alpha * __uint_as_float(r[i]));
                }
                if (out_bf16) {
                    // one slab: 32 rows x 64 columns of bf16 (128 B per row), 128B-swizzled like the TMA box expects
                    const uint32_t slab = slab0 + slab_sel * SLAB_BYTES;
                    if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        st_shared_v4(slab + row_off + ((((uint32_t)q) ^ sw) << 4),
                                     pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                                     pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                                     pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                                     pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (leader) tma_store_2d(&tmC, slab, col0, row0, reduce);
                    slab_sel ^= 1u;
                } else {
                    // two slabs of 32 rows x 32 fp32 columns
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (col0 + 32 * hf < N) {
                            const uint32_t slab = slab0 + slab_sel * SLAB_BYTES;
                            if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            __syncwarp();
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                st_shared_v4(slab + row_off + ((((uint32_t)q) ^ sw) << 4), r[32 * hf + 4 * q], r[32 * hf + 4 * q + 1],
                                             r[32 * hf + 4 * q + 2], r[32 * hf + 4 * q + 3]);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (leader) tma_store_2d(&tmC, slab, col0 + 32 * hf, row0, reduce);
                            slab_sel ^= 1u;
                        }
                    }
                }
            }
            // accumulator fully read into registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (leader) mbar_arrive(tempty_bar(acc));
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}
}  // namespace v2

void encode_map_raw(CUtensorMap* map, const bf16* ptr, int ld, int mn_major, long long rows_mn, long long K, int tile_mn) {
    // K-major  : global (rows_mn, K), K contiguous  -> dims {K, rows_mn}, box {64, tile_mn}
    // MN-major : global (K, rows_mn), rows contiguous -> dims {rows_mn, K}, box {64, 64}
    cuuint64_t dims[2], strides[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    if (!mn_major) {
        dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows_mn;
        box[0] = BK; box[1] = (cuuint32_t)tile_mn;
    } else {
        dims[0] = (cuuint64_t)rows_mn; dims[1] = (cuuint64_t)K;
        box[0] = 64; box[1] = BK;
    }
    strides[0] = (cuuint64_t)ld * sizeof(bf16);
    if (((uintptr_t)ptr & 15) || (strides[0] & 15)) throw std::runtime_error("gemm_tc: operand must be 16-byte aligned with ld % 8 == 0");
    CUresult rc = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
}

// ---- tensor-map cache.  cuTensorMapEncodeTiled costs ~1-2 us on the host and a step re-issues the same few hundred
// (pointer, shape, box) combinations (three encodes per GEMM launch); the per-step generic recurrence of
// wide models is host-bound on them.  Direct-mapped, 1024 entries, keyed on every argument of the encode.
struct MapKey {
    const void* ptr; long long d0, d1, d2; long long s0, s1; int b0, b1, b2, kind;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s0 == o.s0 && s1 == o.s1 && b0 == o.b0 && b1 == o.b1 && b2 == o.b2 && kind == o.kind;
    }
};
struct MapSlot { MapKey key; CUtensorMap map; bool valid = false; };
MapSlot g_map_cache[1024];
std::mutex g_map_mu;
inline size_t map_hash(const MapKey& k) {
    unsigned long long h = (unsigned long long)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long)k.d0 * 0xC2B2AE3D27D4EB4Full + (unsigned long long)k.d1 * 0x165667B19E3779F9ull + (unsigned long long)k.s0 * 31 + k.kind * 7 + k.b1;
    return (size_t)((h >> 20) & 1023);
}
template <class F>
void cached_map(CUtensorMap* out, const MapKey& key, F&& encode) {
    std::lock_guard<std::mutex> lk(g_map_mu);
    MapSlot& sl = g_map_cache[map_hash(key)];
    if (!(sl.valid && sl.key == key)) {
        encode(&sl.map);
        sl.key = key;
        sl.valid = true;
    }
    *out = sl.map;
}

template <int A_MN, int B_MN>
void launch(const CUtensorMap& ta, const CUtensorMap& tb, float* Cf, bf16* Ch, int ldc, int M, int N, int K, float alpha,
            const float* bias, int mode, int kbps, int splits, int vec_ok, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tc<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
    k_gemm_tc<A_MN, B_MN><<<grid, NTHREADS, SMEM_BYTES, s>>>(ta, tb, Cf, Ch, ldc, M, N, K, alpha, bias, mode, kbps, vec_ok);
    COUNT_LAUNCH();
}

void encode_c_map_raw(CUtensorMap* map, void* ptr, int is_bf16, int ldc, long long M, long long N) {
    // output (M, N) row-major; one epilogue warp stores a box of 32 rows x 128 bytes (128B swizzle)
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M}, strides[1] = {(cuuint64_t)ldc * (is_bf16 ? 2 : 4)};
    cuuint32_t box[2] = {is_bf16 ? 64u : 32u, 32u}, estr[2] = {1, 1};
    CUresult rc = g_encode(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (C) failed with code " + std::to_string((int)rc));
}

void encode_map(CUtensorMap* map, const bf16* ptr, int ld, int mn_major, long long rows_mn, long long K, int tile_mn) {
    const MapKey key{ptr, rows_mn, K, 0, ld, 0, tile_mn, mn_major, 0, 1};
    cached_map(map, key, [&](CUtensorMap* m) { encode_map_raw(m, ptr, ld, mn_major, rows_mn, K, tile_mn); });
}
void encode_c_map(CUtensorMap* map, void* ptr, int is_bf16, int ldc, long long M, long long N) {
    const MapKey key{ptr, M, N, 0, ldc, 0, is_bf16, 0, 0, 2};
    cached_map(map, key, [&](CUtensorMap* m) { encode_c_map_raw(m, ptr, is_bf16, ldc, M, N); });
}

thread_local bool g_pdl = false;
template <int A_MN, int B_MN>
void launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, float alpha, const float* bias,
             int out_bf16, int reduce, const v2::Work& W, int grid, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(v2::k_gemm_tc2<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM_BYTES));
        configured = true;
    }
    if (g_pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = v2::SMEM_BYTES; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, v2::k_gemm_tc2<A_MN, B_MN>, ta, tb, tc, M, N, alpha, bias, out_bf16, reduce, W));
    } else {
        v2::k_gemm_tc2<A_MN, B_MN><<<grid, NTHREADS, v2::SMEM_BYTES, s>>>(ta, tb, tc, M, N, alpha, bias, out_bf16, reduce, W);
    }
    COUNT_LAUNCH();
}
}  // namespace
// the next k_gemm_tc2 launches of this thread carry the programmatic-stream-serialization attribute (per-step chains of the
// generic recurrence: the launch latency of step k+1 hides under step k)
void gemm_tc_pdl(bool on) { g_pdl = on; }

// Tensor map over the rows of an activation matrix (rows, ld) bf16 whose rows are dealt round robin over `ns` slices:
// 3D view {x = column, y = row % ns, z = row / ns}, box {64 columns, 1, box_rows} with the 128-byte swizzle -- the rows
// r0, r0 + ns, r0 + 2 ns, ... of one slice arrive as one K-major UMMA operand tile (gru_tc.cu, forward kernel 3)
void tma_encode_slice_rows_bf16(void* map_out, const bf16* base, int ld, long long rows, int ns, int box_rows) {
    if (!g_ready) throw std::runtime_error("TMA descriptors unavailable (gemm_tc_init)");
    if (((uintptr_t)base & 15) || (ld % 8)) throw std::runtime_error("tma_encode_slice_rows: base must be 16-byte aligned with ld % 8 == 0");
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)ns, (cuuint64_t)((rows + ns - 1) / ns)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(bf16), (cuuint64_t)ld * sizeof(bf16) * (cuuint64_t)ns};
    cuuint32_t box[3] = {64u, 1u, (cuuint32_t)box_rows}, estr[3] = {1, 1, 1};
    CUresult rc = g_encode((CUtensorMap*)map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (slice rows) failed with code " + std::to_string((int)rc));
}

// plain 2-D map over a K-major bf16 matrix (rows, ld), box {64 columns, box_rows}, 128-byte swizzle; cached
void tma_encode_2d_bf16(void* map_out, const bf16* base, int ld, long long rows, int cols, int box_rows) {
    if (!g_ready) throw std::runtime_error("TMA descriptors unavailable (gemm_tc_init)");
    encode_map((CUtensorMap*)map_out, base, ld, 0, rows, cols, box_rows);
}

void gemm_tc_init(int device) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (g_tried) return;
    g_tried = true;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return;
    g_num_sms = prop.multiProcessorCount;
    if (prop.major != 10) return;  // tcgen05 / TMEM exist on sm_100 family only
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
        return;
    g_encode = (EncodeTiledFn)fn;
    g_ready = true;
}
bool gemm_tc_available() { return g_ready; }

void gemm_tc(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, float* Cf, bf16* Ch, int ldc, int M, int N,
             int K, float alpha, const float* bias, int accumulate, cudaStream_t s, int short_units) {
    if (!g_ready) throw std::runtime_error("gemm_tc: tcgen05 path not initialised");
    if (M <= 0 || N <= 0) return;
    if (K <= 0) throw std::runtime_error("gemm_tc: K must be positive");
    // ---- v2: persistent 128 x 256 kernel with TMA-store epilogue (one output type, 16-byte aligned rows)
    static const bool force_v1 = getenv("ARGSIM_GEMM_V1") != nullptr;
    const bool one_out = (Cf != nullptr) != (Ch != nullptr);
    const bool c_aligned = Cf ? ((((uintptr_t)Cf & 15) == 0) && ldc % 4 == 0) : ((((uintptr_t)Ch & 15) == 0) && ldc % 8 == 0);
    if (!force_v1 && one_out && c_aligned && !(Ch && accumulate)) {
        v2::Work W;
        W.tiles_m = cdiv(M, v2::BM);
        W.tiles_n = cdiv(N, v2::BN);
        W.nkb_total = cdiv(K, v2::BK);
        const int tiles = W.tiles_m * W.tiles_n;
        int splits = 1;
        if (Cf && tiles * 2 <= g_num_sms && W.nkb_total >= 8) splits = std::max(1, std::min(std::min(W.nkb_total / 4, g_num_sms / tiles), 64));
        // bf16 output with fewer work units than half the SMs and a long K (dho = dlogits . E): two k-halves meet in a bf16
        // reduce-add (0 + p1 is exact, so the result is p1 + p2 rounded once more: deterministic, <= 1 bf16 ulp)
        if (Ch && tiles * 2 <= g_num_sms && W.nkb_total >= 32) splits = 2;
        const bool short_ctas = short_units > 0 && Cf && !bias;
        if (short_ctas) splits = std::max(splits, cdiv(W.nkb_total, short_units));
        W.kb_per_split = cdiv(W.nkb_total, splits);
        W.splits = cdiv(W.nkb_total, W.kb_per_split);   // no empty split
        if (W.splits > 1 && !accumulate) {
            if (Cf) CUDA_CHECK(cudaMemset2DAsync(Cf, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s));
            else CUDA_CHECK(cudaMemset2DAsync(Ch, (size_t)ldc * sizeof(bf16), 0, (size_t)N * sizeof(bf16), M, s));
        }
        const int reduce = (accumulate || W.splits > 1) ? 1 : 0;
        // W.b_first (weight tiles loaded before griddepcontrol.wait) measured neutral on the scaled config (138.9 vs 136.9 ms): the
        // next GEMM's CTAs only become resident when this one's leave, so there is nothing to overlap with; left off
        CUtensorMap ta, tb, tc;
        encode_map(&ta, A, lda, a_mn, M, K, v2::BM);
        encode_map(&tb, B, ldb, b_mn, N, K, v2::BN);
        encode_c_map(&tc, Cf ? (void*)Cf : (void*)Ch, Ch != nullptr, ldc, M, N);
        const int grid = short_ctas ? tiles * W.splits : std::min(tiles * W.splits, g_num_sms);
        const int ob = Ch != nullptr;
        if (!a_mn && !b_mn) launch2<0, 0>(ta, tb, tc, M, N, alpha, bias, ob, reduce, W, grid, s);
        else if (!a_mn && b_mn) launch2<0, 1>(ta, tb, tc, M, N, alpha, bias, ob, reduce, W, grid, s);
        else if (a_mn && !b_mn) launch2<1, 0>(ta, tb, tc, M, N, alpha, bias, ob, reduce, W, grid, s);
        else launch2<1, 1>(ta, tb, tc, M, N, alpha, bias, ob, reduce, W, grid, s);
        return;
    }
    CUtensorMap ta, tb;
    encode_map(&ta, A, lda, a_mn, M, K, BM);
    encode_map(&tb, B, ldb, b_mn, N, K, BN);
    const int tiles = cdiv(N, BN) * cdiv(M, BM);
    const int nkb = cdiv(K, BK);
    int splits = 1;
    if (Cf && !Ch && tiles * 2 <= g_num_sms && nkb >= 8) {
        splits = std::min(std::min(nkb / 4, cdiv(2 * g_num_sms, tiles)), 64);
        if (splits < 1) splits = 1;
    }
    int kbps = cdiv(nkb, splits);
    splits = cdiv(nkb, kbps);  // no empty split
    int mode = accumulate ? EPI_ACCUM : EPI_STORE;
    if (splits > 1) {
        if (!accumulate) CUDA_CHECK(cudaMemset2DAsync(Cf, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s));
        mode = EPI_ATOMIC;
    }
    const int vec_ok = (!Cf || (((uintptr_t)Cf & 15) == 0 && ldc % 4 == 0)) && (!Ch || (((uintptr_t)Ch & 15) == 0 && ldc % 8 == 0));
    if (!a_mn && !b_mn) launch<0, 0>(ta, tb, Cf, Ch, ldc, M, N, K, alpha, bias, mode, kbps, splits, vec_ok, s);
    else if (!a_mn && b_mn) launch<0, 1>(ta, tb, Cf, Ch, ldc, M, N, K, alpha, bias, mode, kbps, splits, vec_ok, s);
    else if (a_mn && !b_mn) launch<1, 0>(ta, tb, Cf, Ch, ldc, M, N, K, alpha, bias, mode, kbps, splits, vec_ok, s);
    else launch<1, 1>(ta, tb, Cf, Ch, ldc, M, N, K, alpha, bias, mode, kbps, splits, vec_ok, s);
}
