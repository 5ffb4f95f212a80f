// gru_tc.cu -- persistent GRU recurrence for H = 512 on the 5th-generation tensor cores (tcgen05, TS form).
//
// Same decomposition as gru_mma.cu (a GROUP of 16 CTAs owns one (direction, batch-slice); CTA c owns hidden units
// [32c, 32c+32) = 96 rows of R), but the recurrent product runs on tcgen05 instead of mma.sync:
//   * A = R_own lives in TENSOR MEMORY for the whole launch: TMEM lane = gate*32 + unit (96 of 128 lanes), the 512 k
//     of a row packed two bf16 per 32-bit column -> 256 of the 512 columns, written once with tcgen05.st;
//   * B = h_{t-1} of the slice's live rows (CN rows x 512 k, bf16) is staged in shared memory in the UMMA K-major
//     128-byte-swizzle layout (8 k-blocks of [CN rows x 128 B]) straight from the exchange buffer;
//   * D[128 x CN] (fp32) accumulates in TMEM: 32 tcgen05.mma (M=128, N=CN, K=16) issued by ONE thread, committed to an
//     mbarrier; the gate warps read it back with tcgen05.ld (thread = one gate row, CN columns = batch rows).
// What that buys over the register-stationary mma.sync kernel: the MMA time of a step is 32 * CN/2 cycles whatever the
// number of live rows (16 rows cost what 8 do), the 8-way k-split reduction through shared memory (24 LDS + adds per
// row) becomes 3 LDS, and a slice can hold 128 rows per step (N = 128) instead of 8 chunks of 16.
// The exchange (all-gather of the new h among the 16 CTAs) is the same in-band-tag LL protocol through L2 as in
// gru_mma.cu: 8-byte words of two bf16 + a 32-bit step tag.  Layout of the exchange buffer here: row-major,
// [parity][row][256 words], so a publishing warp writes one 128-byte line per row and a polling thread reads the 32
// bytes (4 words = 8 units) that make up one 16-byte chunk of the swizzled operand tile.
#include "kernels.h"
#include "plan.h"
#include <cuda.h>
#include <vector>

namespace {

constexpr int HH = 512;
constexpr int CL = 16;     // CTAs per group
constexpr int UN = 32;     // hidden units per CTA
constexpr int NTH = 256;
constexpr int MAX_BSL = 288;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_A = 0;       // columns [0,256): R_own
constexpr uint32_t TM_D = 256;     // columns [256, 256 + CN): accumulator

struct TcDirP {
    const float* gx; const bf16* R; const float* bR; const float* h0;
    float* hs_f; bf16* hs_h; float* cache; float* hT;
    int ld_gx, ld_hs, reverse;
};
struct TcFwdP {
    TcDirP dir[2];
    const int* off; const int* nact;
    unsigned long long* xbuf;
    int ndir, nslices, b, Ttot, t0, Tseg, bslr;
    unsigned tag_base;
    long long* prof;
    int poll_delay;    // cycles a stage warp lets pass after the local publish before its first poll round
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]   (TS form: the A operand is read from tensor memory)
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one lane of a CONVERGED warp; ptxas then issues the tcgen05 instructions of the guarded region from uniform registers
// back to back.  Guarding with `tid == k` instead makes it wrap every tcgen05.mma in an ELECT / BRA.U.ANY loop (the
// operands are not provably warp-uniform): 59 cycles per instruction measured, against 8 for the MMA itself at N = 16.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major tile, 128-byte swizzle (same encoding as gemm_tc.cu make_desc<0>)
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void ll_store(unsigned long long* p, uint32_t data, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const unsigned long long* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_f32(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
__device__ __forceinline__ int slice_rows(int nat, int sl, int ns) { return nat > sl ? (nat - sl + ns - 1) / ns : 0; }
#define POLL_GUARD(t0) if (clock64() - (t0) > 4000000000LL) __trap()
#define NA(t) s_nact[(t) - P.t0 + 1]
#define OFF(t) s_off[(t) - P.t0 + 1]
#define PROF_MARK(i) do { if (P.prof && tid == 0) { const long long now_ = clock64(); pacc[i] += now_ - plast; plast = now_; } } while (0)

// byte offset of the 16-byte chunk (row n, chunk c of 64: 8 units each) inside the operand tile
template <int CN>
__device__ __forceinline__ uint32_t hs_off(int n, int c) {
    return (uint32_t)((c >> 3) * (CN * 128) + (n >> 3) * 1024 + (n & 7) * 128 + (((c & 7) ^ (n & 7)) << 4));
}

// =========================================================================================
// forward.  CN = rows per MMA (N of the instruction): 16, 32, 64 or 128.
// =========================================================================================
template <int CN>
__global__ void __launch_bounds__(NTH, 1) k_gru_tc_fwd(const __grid_constant__ TcFwdP P) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    if ((int)(blockIdx.x / CL) >= P.ndir * P.nslices) return;          // padding CTA: only there for the 128-block placement
    const uint32_t sbase = (smem_u32(sm_raw) + 1023u) & ~1023u;
    unsigned char* const sm = sm_raw + (sbase - smem_u32(sm_raw));
    // [Hs: CN KB][G: 3*CN*32 f32][hst: bslr*32 f32][tables: 2*(Tseg+2) int][bar 8][slot 4]
    const uint32_t Hs = sbase;
    float* G = reinterpret_cast<float*>(sm + CN * 1024);
    float* hst = G + 3 * CN * UN;
    int* s_nact = reinterpret_cast<int*>(hst + (size_t)P.bslr * UN);
    int* s_off = s_nact + P.Tseg + 2;
    unsigned long long* barp = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(s_off + P.Tseg + 2) + 15) & ~uintptr_t(15));
    const uint32_t dbar = smem_u32(barp);
    const uint32_t tslot = dbar + 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = blockIdx.x / CL, c = blockIdx.x % CL;
    const int ns = P.nslices, d = grp / ns, sl = grp % ns;
    const TcDirP& A = P.dir[d];
    for (int i = tid; i < P.Tseg + 2; i += NTH) {
        const int tt = P.t0 - 1 + i;
        s_nact[i] = (tt >= 0 && tt < P.Ttot) ? slice_rows(P.nact[tt], sl, ns) : 0;
        s_off[i] = (tt >= 0 && tt <= P.Ttot) ? P.off[tt] : 0;
    }
    if (tid == 0) {
        mbar_init(dbar, CN <= 32 ? 4 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");

    // ---- R_own -> tensor memory: lane = gate*32 + unit, column j holds k = 2j, 2j+1.  Warps 0..3 own the four lane
    // quadrants; quadrant 3 (lanes 96..127) is unused and zeroed.
    if (warp < 4) {
        const int gate = warp;
        const uint4* src = gate < 3 ? reinterpret_cast<const uint4*>(A.R + (size_t)(gate * HH + UN * c + lane) * HH) : nullptr;
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + TM_A;
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {   // 32 x 8 columns = 256 columns = 512 k
            uint32_t r[8];
            if (src) {
                const uint4 v0 = src[2 * j], v1 = src[2 * j + 1];
                r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = 0u;
            }
            tc_st8(tbase + 8 * j, r);
        }
        tc_wait_st();
    }
    const int col = UN * c + lane;
    const float bRr = A.bR[col], bRu = A.bR[HH + col], bRn = A.bR[2 * HH + col];
    const int nloc = slice_rows(P.b, sl, ns);
    for (int i = tid; i < nloc * UN; i += NTH) {
        const int jl = i >> 5, u = i & 31;
        hst[i] = A.h0 ? A.h0[(size_t)(jl * ns + sl) * HH + UN * c + u] : 0.f;
    }
    const size_t xpar = (size_t)P.bslr * (HH / 2);
    unsigned long long* X = P.xbuf + (size_t)grp * 2 * xpar;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t idesc = make_idesc(CN);
    constexpr int NACC1 = CN <= 32 ? 4 : 1;
    int na_prev = 0;
    bool first = true;
    uint32_t dphase = 0;
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long plast = clock64();
    constexpr int NCHUNK16 = CN * 64;            // 16-byte operand chunks per MMA chunk
    constexpr int PER_T = NCHUNK16 / NTH;        // per thread: CN / 4
    constexpr int U = PER_T < 8 ? PER_T : 8;     // chunks polled together (2 x ld.v4 each)
    constexpr int RPT = CN / 8;                  // rows per thread in the gate phase
    for (int k = 0; k < P.Tseg; ++k) {
        const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
        const int na = NA(t);
        if (na == 0) {
            if (A.reverse) continue;
            break;
        }
        PROF_MARK(0);
        const long long row_base = OFF(t);
        const int npoll = first ? 0 : min(na, na_prev);
        const unsigned tag = P.tag_base + (unsigned)(k - 1), tagw = P.tag_base + (unsigned)k;
        const unsigned long long* Xr = X + (size_t)((k - 1) & 1) * xpar;
        unsigned long long* Xw = X + (size_t)(k & 1) * xpar;
        for (int ch = 0; ch * CN < na; ++ch) {
            const int nrows = min(CN, na - ch * CN);
            // ---------------- operand staging: h_{t-1} of the chunk's rows -> swizzled shared memory
#pragma unroll 1
            for (int i0 = 0; i0 < PER_T; i0 += U) {
                uint4 x[U][2];
                bool need[U];
                int nn[U], cc[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = (i0 + u) * NTH + tid;
                    nn[u] = idx >> 6; cc[u] = idx & 63;
                    need[u] = nn[u] < nrows && ch * CN + nn[u] < npoll;
                }
                if (!first) {
                    bool ok;
                    const long long tp0 = clock64();
                    do {
                        ok = true;
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            if (need[u]) {
                                const unsigned long long* src = Xr + ((size_t)(ch * CN + nn[u]) * (HH / 2) + 4 * cc[u]);
                                x[u][0] = ll_load2(src);
                                x[u][1] = ll_load2(src + 2);
                            }
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            if (need[u] && (x[u][0].y != tag || x[u][0].w != tag || x[u][1].y != tag || x[u][1].w != tag)) ok = false;
                        POLL_GUARD(tp0);
                    } while (!ok);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (nn[u] >= nrows) continue;
                    uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
                    if (need[u]) {
                        w0 = x[u][0].x; w1 = x[u][0].z; w2 = x[u][1].x; w3 = x[u][1].z;
                    } else if (first && A.h0) {
                        const float4* hp = reinterpret_cast<const float4*>(A.h0 + (size_t)((ch * CN + nn[u]) * ns + sl) * HH + 8 * cc[u]);
                        const float4 a0 = hp[0], a1 = hp[1];
                        w0 = pack2(a0.x, a0.y); w1 = pack2(a0.z, a0.w); w2 = pack2(a1.x, a1.y); w3 = pack2(a1.z, a1.w);
                    }
                    st_shared_v4(Hs + hs_off<CN>(nn[u], cc[u]), w0, w1, w2, w3);
                }
            }
            PROF_MARK(1);   // poll + staging
            // gx of the rows this thread finishes: L2 hits (prefetched two steps ago), complete under the MMAs
            float gxv[RPT][3];
#pragma unroll
            for (int e = 0; e < RPT; ++e) {
                const int n = warp + 8 * e;
                gxv[e][0] = gxv[e][1] = gxv[e][2] = 0.f;
                if (n < nrows) {
                    const float* gp = A.gx + (size_t)(row_base + (long long)(ch * CN + n) * ns + sl) * A.ld_gx + col;
                    gxv[e][0] = ld_f32(gp); gxv[e][1] = ld_f32(gp + HH); gxv[e][2] = ld_f32(gp + 2 * HH);
                }
            }
            fence_async_smem();
            __syncthreads();
            PROF_MARK(2);   // barrier
            // ---------------- D[128 x CN] = R_own[128 x 512] . h^T.  CN <= 32: warps 0..3 issue 8 MMAs each (their own quarter
            // of K) into their own accumulator -- one thread needs ~19 cycles per tcgen05.mma, the tensor pipe 8 at N = 16;
            // the gate warps add the four partial sums.  CN >= 64: tensor-bound, one issuer (warp 3), one accumulator.
            if (NACC1 == 4) {
                if (warp < 4) {
                    if (elect_one()) {
                        tc_fence_after();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int jj = 8 * warp + j;
                            const uint64_t db = make_desc_k(Hs + (uint32_t)((jj >> 2) * (CN * 128) + (jj & 3) * 32));
                            tc_mma_ts(tmem + TM_D + (uint32_t)(warp * CN), tmem + TM_A + 8 * jj, db, idesc, j > 0 ? 1u : 0u);
                        }
                        tc_commit(dbar);
                    }
                    __syncwarp();
                }
            } else if (warp == 3) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint64_t db = make_desc_k(Hs + (uint32_t)((j >> 2) * (CN * 128) + (j & 3) * 32));
                        tc_mma_ts(tmem + TM_D, tmem + TM_A + 8 * j, db, idesc, j > 0 ? 1u : 0u);
                    }
                    tc_commit(dbar);
                }
                __syncwarp();
            }
            // L2 prefetch of the gx rows of the step after next (one warp per row, round robin)
            if (k + 2 < P.Tseg && ch == 0) {
                const int tn = A.reverse ? t - 2 : t + 2;
                const int nan = NA(tn);
                const long long rbn = OFF(tn);
                for (int n = warp; n < nan; n += NTH / 32) {
                    const float* gp = A.gx + (size_t)(rbn + (long long)n * ns + sl) * A.ld_gx + UN * c;
                    if (lane < 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + lane * HH));
                }
            }
            // ---------------- accumulator -> registers -> G[gate][row][unit]: warps 0-2 read columns [0, CN/2), warps 4-6
            // columns [CN/2, CN) of their lane quadrant (= gate)
            if ((warp & 3) < 3) {
                mbar_wait(dbar, dphase);
                tc_fence_after();
                const int gate = warp & 3, half = warp >> 2;
                const uint32_t ta = tmem + ((uint32_t)(gate * 32) << 16) + TM_D + (uint32_t)(half * (CN / 2));
                uint32_t r[NACC1][CN / 2];
#pragma unroll
                for (int a = 0; a < NACC1; ++a)
#pragma unroll
                    for (int i = 0; i < CN / 16; ++i) tc_ld8(ta + (uint32_t)(a * CN) + 8 * i, r[a] + 8 * i);
                tc_wait_ld();
                float* gp = G + ((size_t)gate * CN + half * (CN / 2)) * UN + lane;
#pragma unroll
                for (int i = 0; i < CN / 2; ++i) {
                    float v = __uint_as_float(r[0][i]);
#pragma unroll
                    for (int a = 1; a < NACC1; ++a) v += __uint_as_float(r[a][i]);
                    gp[i * UN] = v;
                }
                tc_fence_before();
            }
            dphase ^= 1u;
            PROF_MARK(3);   // MMA + accumulator read
            __syncthreads();
            PROF_MARK(4);
            // ---------------- gates: lane = unit, warp -> rows w, w+8, ...
#pragma unroll
            for (int e = 0; e < RPT; ++e) {
                const int n = warp + 8 * e;
                if (n < nrows) {
                    const float s0 = G[(0 * CN + n) * UN + lane], s1 = G[(1 * CN + n) * UN + lane], s2 = G[(2 * CN + n) * UN + lane];
                    const float r = sigm(gxv[e][0] + s0 + bRr);
                    const float z = sigm(gxv[e][1] + s1 + bRu);
                    const float qq = s2 + bRn;
                    const float nv = tanh_fast(gxv[e][2] + r * qq);
                    const int jl = ch * CN + n;
                    const float hp = hst[jl * UN + lane];
                    const float h = (1.f - z) * nv + z * hp;
                    const __nv_bfloat16 hb16 = __float2bfloat16(h);
                    // publish first: this store is on every peer's critical path
                    const uint32_t hb = (uint32_t)__bfloat16_as_ushort(hb16);
                    const uint32_t ob = __shfl_down_sync(0xffffffffu, hb, 1);
                    if (!(lane & 1)) ll_store(Xw + (size_t)jl * (HH / 2) + (col >> 1), hb | (ob << 16), tagw);
                    hst[jl * UN + lane] = h;
                    const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                    if (A.hs_h) A.hs_h[row * A.ld_hs + col] = hb16;
                    if (A.hs_f) A.hs_f[row * A.ld_hs + col] = h;
                    if (A.cache) {
                        float* cp = A.cache + row * 4 * HH + col;
                        cp[0] = r; cp[HH] = z; cp[2 * HH] = nv; cp[3 * HH] = qq;
                    }
                }
            }
            PROF_MARK(5);   // gates + publish + bookkeeping
        }
        na_prev = na;
        first = false;
        PROF_MARK(6);
    }
    __syncthreads();
    if (A.hT) {   // state handed to the next time segment of this layer
        for (int i = tid; i < nloc * UN; i += NTH) {
            const int jl = i >> 5, u = i & 31;
            A.hT[(size_t)(jl * ns + sl) * HH + UN * c + u] = hst[i];
        }
    }
    if (P.prof && tid == 0)
        for (int i = 0; i < 8; ++i) P.prof[blockIdx.x * 8 + i] = pacc[i];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}


// =========================================================================================
// forward, warp-specialised (the default): three roles that hand chunks of CN rows to each other through mbarriers
//   warps 8..11  STAGE : warp w polls k-blocks 2(w-8), 2(w-8)+1 (128 units = the slices of 4 CTAs) of the chunk's rows
//                        from the exchange buffer and writes them into the operand tile (slot = chunk number & 1);
//   warp  12     MMA   : as each pair of k-blocks lands, one lane issues its 8 tcgen05.mma (K = 16 each) into the slot's
//                        accumulator; after the 32nd it commits to the slot's "accumulator full" mbarrier;
//   warps 0..7   GATE  : read the accumulator (lane quadrant = gate, two column halves), combine the three gates of a
//                        unit through shared memory, finish the GRU cell, publish the new h, write the bookkeeping.
// No block-wide barrier in the loop.  With one chunk per step (few live rows: the latency regime) the roles simply take
// turns and the MMAs of a k-block start while the other k-blocks are still in flight; with several chunks per step
// (many rows) the staging of chunk i+1 and the gate math of chunk i-1 run under the MMAs of chunk i.
// CN = 16, 32 or 64 (two operand slots of CN KB, two accumulators of CN columns).
// =========================================================================================
constexpr int NTH2 = 416;    // 13 warps: 8 gate, 4 stage, 1 MMA
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gate_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int CN>
__global__ void __launch_bounds__(NTH2, 1) k_gru_tc_fwd2(const __grid_constant__ TcFwdP P) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    if ((int)(blockIdx.x / CL) >= P.ndir * P.nslices) return;
    const uint32_t sbase = (smem_u32(sm_raw) + 1023u) & ~1023u;
    unsigned char* const sm = sm_raw + (sbase - smem_u32(sm_raw));
    // [Hs: 2 x CN KB][G: 2 x 3*CN*32 f32][hst: bslr*32 f32][tables: 2*(Tseg+2) int][barriers: 12 x 8 B][slot 4]
    const uint32_t Hs0 = sbase;
    float* G0 = reinterpret_cast<float*>(sm + 2 * CN * 1024);
    float* hst = G0 + 2 * 3 * CN * UN;
    int* s_nact = reinterpret_cast<int*>(hst + (size_t)P.bslr * UN);
    int* s_off = s_nact + P.Tseg + 2;
    unsigned long long* barp = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(s_off + P.Tseg + 2) + 15) & ~uintptr_t(15));
    const uint32_t bars = smem_u32(barp);
    auto kfull = [&](int slot, int kp) { return bars + 8u * (uint32_t)(slot * 4 + kp); };
    auto dfull = [&](int slot) { return bars + 8u * (uint32_t)(8 + slot); };
    auto dfree = [&](int slot) { return bars + 8u * (uint32_t)(10 + slot); };
    const uint32_t tslot = bars + 8u * 12;
    // chunks whose new h this CTA's own gate warps have published (8 increments per chunk): the stage warps start to
    // poll a chunk only when its predecessor has been published HERE -- the peers run in lock step, so their words are
    // then in flight; polling any earlier only loads the L2 (measured: 3,400 instead of 1,000 cycles until the data shows)
    volatile int* s_pub = reinterpret_cast<volatile int*>(barp + 13);
    // CN <= 32: every stage warp issues the 8 MMAs of its own k-blocks into its own accumulator (4 per slot, summed by the
    // gate warps): the issue of 32 MMAs (~19 cycles each from one thread) is spread over 4 threads.  CN = 64: the MMAs
    // are tensor-bound (32 cycles each), one accumulator per slot, issued by the MMA warp.
    constexpr bool SPLIT = CN <= 32;
    constexpr int NACC = SPLIT ? 4 : 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = blockIdx.x / CL, c = blockIdx.x % CL;
    const int ns = P.nslices, d = grp / ns, sl = grp % ns;
    const TcDirP& A = P.dir[d];
    if (tid == 0) *s_pub = 0;
    for (int i = tid; i < P.Tseg + 2; i += NTH2) {
        const int tt = P.t0 - 1 + i;
        s_nact[i] = (tt >= 0 && tt < P.Ttot) ? slice_rows(P.nact[tt], sl, ns) : 0;
        s_off[i] = (tt >= 0 && tt <= P.Ttot) ? P.off[tt] : 0;
    }
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bars + 8u * i, 1);
        mbar_init(dfull(0), SPLIT ? 4 : 1); mbar_init(dfull(1), SPLIT ? 4 : 1);
        mbar_init(dfree(0), 6); mbar_init(dfree(1), 6);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");

    if (warp < 4) {   // R_own -> tensor memory (see k_gru_tc_fwd)
        const int gate = warp;
        const uint4* src = gate < 3 ? reinterpret_cast<const uint4*>(A.R + (size_t)(gate * HH + UN * c + lane) * HH) : nullptr;
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + TM_A;
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
            uint32_t r[8];
            if (src) {
                const uint4 v0 = src[2 * j], v1 = src[2 * j + 1];
                r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = 0u;
            }
            tc_st8(tbase + 8 * j, r);
        }
        tc_wait_st();
    }
    const int nloc = slice_rows(P.b, sl, ns);
    for (int i = tid; i < nloc * UN; i += NTH2) {
        const int jl = i >> 5, u = i & 31;
        hst[i] = A.h0 ? A.h0[(size_t)(jl * ns + sl) * HH + UN * c + u] : 0.f;
    }
    const size_t xpar = (size_t)P.bslr * (HH / 2);
    unsigned long long* X = P.xbuf + (size_t)grp * 2 * xpar;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long plast = clock64();
#define PROF_MARK_L(i) do { if (P.prof && lane == 0) { const long long now_ = clock64(); pacc[i] += now_ - plast; plast = now_; } } while (0)
    if (warp < 8) {
        // =========================== GATE warps ===========================
        const int gw = warp, quad = warp & 3, half = warp >> 2;
        const int col = UN * c + lane;
        const float bRr = A.bR[col], bRu = A.bR[HH + col], bRn = A.bR[2 * HH + col];
        constexpr int RPT = CN / 8;
        int q = 0;
        for (int k = 0; k < P.Tseg; ++k) {
            const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
            const int na = NA(t);
            if (na == 0) {
                if (A.reverse) continue;
                break;
            }
            const long long row_base = OFF(t);
            const unsigned tagw = P.tag_base + (unsigned)k;
            unsigned long long* Xw = X + (size_t)(k & 1) * xpar;
            for (int ch = 0; ch * CN < na; ++ch, ++q) {
                const int slot = q & 1, u = q >> 1;
                const int nrows = min(CN, na - ch * CN);
                PROF_MARK(0);
                float gxv[RPT][3];
#pragma unroll
                for (int e = 0; e < RPT; ++e) {
                    const int n = gw + 8 * e;
                    gxv[e][0] = gxv[e][1] = gxv[e][2] = 0.f;
                    if (n < nrows) {
                        const float* gp = A.gx + (size_t)(row_base + (long long)(ch * CN + n) * ns + sl) * A.ld_gx + col;
                        gxv[e][0] = ld_f32(gp); gxv[e][1] = ld_f32(gp + HH); gxv[e][2] = ld_f32(gp + 2 * HH);
                    }
                }
                float* G = G0 + (size_t)slot * 3 * CN * UN;
                if (quad < 3) {
                    mbar_wait(dfull(slot), (uint32_t)(u & 1));
                    tc_fence_after();
                    PROF_MARK(1);   // waiting for the accumulator: exchange + staging + MMA
                    const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + TM_D + (uint32_t)(slot * NACC * CN + half * (CN / 2));
                    uint32_t r[NACC][CN / 2];
#pragma unroll
                    for (int a = 0; a < NACC; ++a)
#pragma unroll
                        for (int i = 0; i < CN / 16; ++i) tc_ld8(ta + (uint32_t)(a * CN) + 8 * i, r[a] + 8 * i);
                    tc_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dfree(slot));
                    float* gp = G + ((size_t)quad * CN + half * (CN / 2)) * UN + lane;
#pragma unroll
                    for (int i = 0; i < CN / 2; ++i) {
                        float v = __uint_as_float(r[0][i]);
#pragma unroll
                        for (int a = 1; a < NACC; ++a) v += __uint_as_float(r[a][i]);
                        gp[i * UN] = v;
                    }
                }
                gate_bar();
                PROF_MARK(2);   // accumulator read + gate exchange through shared memory
#pragma unroll
                for (int e = 0; e < RPT; ++e) {
                    const int n = gw + 8 * e;
                    if (n < nrows) {
                        const float s0 = G[(0 * CN + n) * UN + lane], s1 = G[(1 * CN + n) * UN + lane], s2 = G[(2 * CN + n) * UN + lane];
                        const float r = sigm(gxv[e][0] + s0 + bRr);
                        const float z = sigm(gxv[e][1] + s1 + bRu);
                        const float qq = s2 + bRn;
                        const float nv = tanh_fast(gxv[e][2] + r * qq);
                        const int jl = ch * CN + n;
                        const float hp = hst[jl * UN + lane];
                        const float h = (1.f - z) * nv + z * hp;
                        const __nv_bfloat16 hb16 = __float2bfloat16(h);
                        const uint32_t hb = (uint32_t)__bfloat16_as_ushort(hb16);
                        const uint32_t ob = __shfl_down_sync(0xffffffffu, hb, 1);
                        if (!(lane & 1)) ll_store(Xw + (size_t)jl * (HH / 2) + (col >> 1), hb | (ob << 16), tagw);
                        hst[jl * UN + lane] = h;
                        const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                        if (A.hs_h) A.hs_h[row * A.ld_hs + col] = hb16;
                        if (A.hs_f) A.hs_f[row * A.ld_hs + col] = h;
                        if (A.cache) {
                            float* cp = A.cache + row * 4 * HH + col;
                            cp[0] = r; cp[HH] = z; cp[2 * HH] = nv; cp[3 * HH] = qq;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) atomicAdd(const_cast<int*>(s_pub), 1);
                PROF_MARK(3);   // gates + publish + bookkeeping
                if (k + 2 < P.Tseg && ch == 0) {   // gx rows of the step after next -> L2
                    const int tn = A.reverse ? t - 2 : t + 2;
                    const int nan = NA(tn);
                    const long long rbn = OFF(tn);
                    for (int n = gw; n < nan; n += 8) {
                        const float* gp = A.gx + (size_t)(rbn + (long long)n * ns + sl) * A.ld_gx + UN * c;
                        if (lane < 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + lane * HH));
                    }
                }
                PROF_MARK(4);
            }
        }
    } else if (warp < 12) {
        // =========================== STAGE warps ===========================
        const int kp = warp - 8;                       // pair of k-blocks 2 kp, 2 kp + 1: units [128 kp, 128 kp + 128)
        constexpr int U = CN < 16 ? CN : 16;           // rows polled together (one ld.v4 per row and lane)
        const uint32_t idesc = make_idesc(CN);
        int q = 0, na_prev = 0, q_prev_step = 0;
        bool first = true;
        for (int k = 0; k < P.Tseg; ++k) {
            const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
            const int na = NA(t);
            if (na == 0) {
                if (A.reverse) continue;
                break;
            }
            const int npoll = first ? 0 : min(na, na_prev);
            const unsigned tag = P.tag_base + (unsigned)(k - 1);
            const unsigned long long* Xr = X + (size_t)((k - 1) & 1) * xpar;
            const int q_this_step = q;
            for (int ch = 0; ch * CN < na; ++ch, ++q) {
                const int slot = q & 1, u = q >> 1;
                const int nrows = min(CN, na - ch * CN);
                const uint32_t Hs = Hs0 + (uint32_t)(slot * CN * 1024);
                PROF_MARK_L(0);
                if (u >= 1) mbar_wait(dfull(slot), (uint32_t)((u - 1) & 1));   // the MMAs that read this slot last are complete
                if (ch * CN < npoll) {   // the predecessor chunk (previous step, same rows) has been published by this CTA
                    const int want = 8 * (q_prev_step + ch + 1);
                    const long long tw0 = clock64();
                    bool waited = false;
                    while (*s_pub < want) { waited = true; POLL_GUARD(tw0); }
                    // the peers publish within a few hundred cycles of this CTA and their words need an L2 round trip to show:
                    // a poll round issued right away comes back empty and the next one costs a full round (~700 cycles
                    // chip-wide for 32 KB of LL words per SM); one round issued a little later finds everything
                    if (waited && P.poll_delay > 0) {
                        const long long td0 = clock64();
                        while (clock64() - td0 < P.poll_delay) {}
                    }
                }
                PROF_MARK_L(1);   // stage: waiting for the slot and for the local publish
                // one ld.v4 per row: the warp reads the 512 contiguous bytes (64 LL words = 128 units) of this k-block pair,
                // lane l the words of units 4l .. 4l+3 -> every load instruction is 4 full lines, nothing is fetched twice
#pragma unroll 1
                for (int r0 = 0; r0 < CN; r0 += U) {
                    if (r0 >= nrows) break;
                    uint4 x[U];
                    bool miss[U];
#pragma unroll
                    for (int uu = 0; uu < U; ++uu) miss[uu] = !first && (r0 + uu) < nrows && ch * CN + r0 + uu < npoll;   // warp-uniform
                    if (!first) {
                        bool ok;
                        const long long tp0 = clock64();
                        do {
                            ok = true;
                            if (P.prof && lane == 0) pacc[4] += 1;    // poll rounds
#pragma unroll
                            for (int uu = 0; uu < U; ++uu)
                                if (miss[uu]) x[uu] = ll_load2(Xr + ((size_t)(ch * CN + r0 + uu) * (HH / 2) + 64 * kp + 2 * lane));
#pragma unroll
                            for (int uu = 0; uu < U; ++uu)
                                if (miss[uu]) {
                                    const bool m = __any_sync(0xffffffffu, x[uu].y != tag || x[uu].w != tag);
                                    if (m) ok = false;
                                    else {   // the row is complete: operand store, then it is not polled again
                                        miss[uu] = false;
                                        const uint32_t a = Hs + hs_off<CN>(r0 + uu, 16 * kp + (lane >> 1)) + ((lane & 1) << 3);
                                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x[uu].x), "r"(x[uu].z) : "memory");
                                    }
                                }
                            POLL_GUARD(tp0);
                        } while (!ok);
                    }
                    // rows that are not polled: initial state (first step) or zeros (rows that join here, reverse direction)
#pragma unroll
                    for (int uu = 0; uu < U; ++uu) {
                        const int r = r0 + uu;
                        if (r >= nrows || (!first && ch * CN + r < npoll)) continue;
                        uint32_t w0 = 0u, w1 = 0u;
                        if (first && A.h0) {
                            const float4 a0 = *reinterpret_cast<const float4*>(A.h0 + (size_t)((ch * CN + r) * ns + sl) * HH + 128 * kp + 4 * lane);
                            w0 = pack2(a0.x, a0.y); w1 = pack2(a0.z, a0.w);
                        }
                        const uint32_t a = Hs + hs_off<CN>(r, 16 * kp + (lane >> 1)) + ((lane & 1) << 3);
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(w0), "r"(w1) : "memory");
                    }
                }
                PROF_MARK_L(2);   // stage: poll + operand stores
                fence_async_smem();
                __syncwarp();
                if (SPLIT) {
                    if (u >= 1) {   // the gate warps have read the previous use of this slot's accumulators
                        mbar_wait(dfree(slot), (uint32_t)((u - 1) & 1));
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t td = tmem + TM_D + (uint32_t)((slot * NACC + kp) * CN);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint64_t db = make_desc_k(Hs + (uint32_t)((2 * kp + (j >> 2)) * (CN * 128) + (j & 3) * 32));
                            tc_mma_ts(td, tmem + TM_A + 8 * (8 * kp + j), db, idesc, j > 0 ? 1u : 0u);
                        }
                        tc_commit(dfull(slot));
                    }
                    __syncwarp();
                } else if (lane == 0) {
                    mbar_arrive(kfull(slot, kp));
                }
                PROF_MARK_L(3);
            }
            q_prev_step = q_this_step;
            na_prev = na;
            first = false;
        }
        if (P.prof && warp == 8 && lane == 0)
            for (int i = 0; i < 5; ++i) P.prof[(size_t)blockIdx.x * 16 + 8 + i] = pacc[i];
    } else if (!SPLIT) {
        // =========================== MMA warp ===========================
        const uint32_t idesc = make_idesc(CN);
        int q = 0;
        for (int k = 0; k < P.Tseg; ++k) {
            const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
            const int na = NA(t);
            if (na == 0) {
                if (A.reverse) continue;
                break;
            }
            for (int ch = 0; ch * CN < na; ++ch, ++q) {
                const int slot = q & 1, u = q >> 1;
                const uint32_t Hs = Hs0 + (uint32_t)(slot * CN * 1024);
                const uint32_t td = tmem + TM_D + (uint32_t)(slot * CN);
                PROF_MARK_L(0);
                if (u >= 1) {   // the gate warps have read the previous use of this accumulator
                    mbar_wait(dfree(slot), (uint32_t)((u - 1) & 1));
                    tc_fence_after();
                }
                PROF_MARK_L(1);   // MMA warp: waiting for the accumulator to be free
#pragma unroll
                for (int p4 = 0; p4 < 4; ++p4) {
                    mbar_wait(kfull(slot, p4), (uint32_t)(u & 1));
                    tc_fence_after();
                    if (p4 == 0) PROF_MARK_L(2);   // ... for the first pair of k-blocks
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint64_t db = make_desc_k(Hs + (uint32_t)((2 * p4 + (j >> 2)) * (CN * 128) + (j & 3) * 32));
                            tc_mma_ts(td, tmem + TM_A + 8 * (8 * p4 + j), db, idesc, (p4 > 0 || j > 0) ? 1u : 0u);
                        }
                        if (p4 == 3) tc_commit(dfull(slot));
                    }
                    __syncwarp();
                }
                PROF_MARK_L(3);   // ... for the other three pairs + issue
            }
        }
        if (P.prof && lane == 0)
            for (int i = 0; i < 4; ++i) P.prof[(size_t)blockIdx.x * 16 + 12 + i] = pacc[i];
    }
    tc_fence_before();
    __syncthreads();
    if (A.hT) {
        for (int i = tid; i < nloc * UN; i += NTH2) {
            const int jl = i >> 5, u = i & 31;
            A.hT[(size_t)(jl * ns + sl) * HH + UN * c + u] = hst[i];
        }
    }
    if (P.prof && tid == 0)
        for (int i = 0; i < 8; ++i) P.prof[(size_t)blockIdx.x * 16 + i] = pacc[i];
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// =========================================================================================
// forward, throughput form (many rows per slice): the exchange is NOT a polled LL buffer.  The new h a gate warp writes
// to the layer's output matrix hs (bf16) IS the published value; a CTA raises a flag (release) per (chunk, step) once
// its 32 columns of the chunk's rows are written, the producer warp of every CTA waits for the 16 flags of the
// predecessor chunk (acquire) and then lets the TMA engine copy the chunk's rows of h_{t-1} -- all 512 columns, straight
// from hs through a 3-D tensor map {column, row % ns, row / ns} -- into the swizzled operand tile: 1 KB per row instead
// of 2 KB of tagged words, no load/store instructions, no registers.  Roles: warps 0..7 gate, warp 8 producer (flags +
// TMA), warp 9 MMA; two operand slots and two accumulators of CN = 64 rows, so the copy of chunk i+1, the 32 MMAs
// (N = 64, tensor-bound) of chunk i and the gate math of chunk i-1 overlap.  Latency per chunk is worse than the LL
// exchange (a release fence and two L2 round trips); with >= 2 chunks per step in flight that is hidden.
// Used when no initial / final state is handed over (h0 == hT == null: whole-layer launches).
// =========================================================================================
// threads: GW gate warps + producer + MMA warp
struct TcFwd3X {
    unsigned* flags;     // [group][CL][MAXCH3]
    int xcol[2];         // column of the direction's first unit in the hs matrix the tensor map covers
};
constexpr int MAXCH3 = 8;
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
template <int CN, int NS, int GW>
__global__ void __launch_bounds__((GW + 2) * 32, 1) k_gru_tc_fwd3(const __grid_constant__ TcFwdP P, const __grid_constant__ CUtensorMap tmH,
                                                         const __grid_constant__ TcFwd3X XP) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    if ((int)(blockIdx.x / CL) >= P.ndir * P.nslices) return;
    const uint32_t sbase = (smem_u32(sm_raw) + 1023u) & ~1023u;
    unsigned char* const sm = sm_raw + (sbase - smem_u32(sm_raw));
    // NS operand slots / accumulators of CN rows each; ONE gate-exchange buffer (two gate barriers per chunk protect it)
    // [Hs: NS x CN KB][G: 3*CN*32 f32][hst: bslr*32 f32][tables][barriers: full[NS], dfull[NS], dfree[NS]][slot]
    const uint32_t Hs0 = sbase;
    constexpr int NTH3 = (GW + 2) * 32;
    constexpr int NPART = GW / 4;            // column parts: the GW/4 warps of a lane quadrant split the CN accumulator columns
    constexpr int PC = CN / NPART;           // columns per part: a multiple of 4
    static_assert(PC % 4 == 0 && CN % GW == 0, "chunk / gate-warp geometry");
    float* G0 = reinterpret_cast<float*>(sm + NS * CN * 1024);
    float* hst = G0 + 3 * CN * UN;
    int* s_nact = reinterpret_cast<int*>(hst + (size_t)P.bslr * UN);
    int* s_off = s_nact + P.Tseg + 2;
    unsigned long long* barp = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(s_off + P.Tseg + 2) + 15) & ~uintptr_t(15));
    const uint32_t bars = smem_u32(barp);
    auto full = [&](int slot) { return bars + 8u * (uint32_t)slot; };
    auto dfull = [&](int slot) { return bars + 8u * (uint32_t)(NS + slot); };
    auto dfree = [&](int slot) { return bars + 8u * (uint32_t)(2 * NS + slot); };
    const uint32_t tslot = bars + 8u * (3 * NS);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = blockIdx.x / CL, c = blockIdx.x % CL;
    const int ns = P.nslices, d = grp / ns, sl = grp % ns;
    const TcDirP& A = P.dir[d];
    for (int i = tid; i < P.Tseg + 2; i += NTH3) {
        const int tt = P.t0 - 1 + i;
        s_nact[i] = (tt >= 0 && tt < P.Ttot) ? slice_rows(P.nact[tt], sl, ns) : 0;
        s_off[i] = (tt >= 0 && tt <= P.Ttot) ? P.off[tt] : 0;
    }
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(full(i), 1); mbar_init(dfull(i), 1); mbar_init(dfree(i), 3 * NPART); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmH) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");
    if (warp < 4) {   // R_own -> tensor memory (see k_gru_tc_fwd)
        const int gate = warp;
        const uint4* src = gate < 3 ? reinterpret_cast<const uint4*>(A.R + (size_t)(gate * HH + UN * c + lane) * HH) : nullptr;
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + TM_A;
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
            uint32_t r[8];
            if (src) {
                const uint4 v0 = src[2 * j], v1 = src[2 * j + 1];
                r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = 0u;
            }
            tc_st8(tbase + 8 * j, r);
        }
        tc_wait_st();
    }
    const int nloc = slice_rows(P.b, sl, ns);
    for (int i = tid; i < nloc * UN; i += NTH3) hst[i] = 0.f;
    unsigned* const flags = XP.flags + (size_t)grp * CL * MAXCH3;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < GW) {
        // =========================== GATE warps ===========================
        const int gw = warp, quad = warp & 3, part = warp >> 2;
        const int col = UN * c + lane;
        const float bRr = A.bR[col], bRu = A.bR[HH + col], bRn = A.bR[2 * HH + col];
        constexpr int RPT = CN / GW;
        int q = 0, na_prev = 0, a = 0;
        for (int k = 0; k < P.Tseg; ++k) {
            const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
            const int na = NA(t);
            if (na == 0) {
                if (A.reverse) continue;
                break;
            }
            const long long row_base = OFF(t);
            const int npoll = a == 0 ? 0 : min(na, na_prev);
            for (int ch = 0; ch * CN < na; ++ch, ++q) {
                const int slot = q % NS, u = q / NS;
                const int nrows = min(CN, na - ch * CN);
                float gxv[RPT][3];
#pragma unroll
                for (int e = 0; e < RPT; ++e) {
                    const int n = gw + GW * e;
                    gxv[e][0] = gxv[e][1] = gxv[e][2] = 0.f;
                    if (n < nrows) {
                        const float* gp = A.gx + (size_t)(row_base + (long long)(ch * CN + n) * ns + sl) * A.ld_gx + col;
                        gxv[e][0] = ld_f32(gp); gxv[e][1] = ld_f32(gp + HH); gxv[e][2] = ld_f32(gp + 2 * HH);
                    }
                }
                float* G = G0;
                if (quad < 3) {
                    mbar_wait(dfull(slot), (uint32_t)(u & 1));
                    tc_fence_after();
                    const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + TM_D + (uint32_t)(slot * CN + part * PC);
                    uint32_t r[PC];
#pragma unroll
                    for (int i = 0; i < PC / 8; ++i) tc_ld8(ta + 8 * i, r + 8 * i);
                    if (PC % 8) tc_ld4(ta + (PC / 8) * 8, r + (PC / 8) * 8);
                    tc_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dfree(slot));
                    float* gp = G + ((size_t)quad * CN + part * PC) * UN + lane;
#pragma unroll
                    for (int i = 0; i < PC; ++i) gp[i * UN] = __uint_as_float(r[i]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(GW * 32) : "memory");
#pragma unroll
                for (int e = 0; e < RPT; ++e) {
                    const int n = gw + GW * e;
                    if (n < nrows) {
                        const int jl = ch * CN + n;
                        // rows without a predecessor (first step; rows that join here) start from h = 0: R h = 0, whatever the
                        // copy engine brought into their operand rows
                        const bool live = jl < npoll;
                        const float s0 = live ? G[(0 * CN + n) * UN + lane] : 0.f, s1 = live ? G[(1 * CN + n) * UN + lane] : 0.f,
                                    s2 = live ? G[(2 * CN + n) * UN + lane] : 0.f;
                        const float r = sigm(gxv[e][0] + s0 + bRr);
                        const float z = sigm(gxv[e][1] + s1 + bRu);
                        const float qq = s2 + bRn;
                        const float nv = tanh_fast(gxv[e][2] + r * qq);
                        const float hp = hst[jl * UN + lane];
                        const float h = (1.f - z) * nv + z * hp;
                        hst[jl * UN + lane] = h;
                        const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                        A.hs_h[row * A.ld_hs + col] = __float2bfloat16(h);     // the publish
                        if (A.hs_f) A.hs_f[row * A.ld_hs + col] = h;
                        if (A.cache) {
                            float* cp = A.cache + row * 4 * HH + col;
                            cp[0] = r; cp[HH] = z; cp[2 * HH] = nv; cp[3 * HH] = qq;
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(GW * 32) : "memory");   // all gate warps have written their rows of the chunk (and are done with G)
                if (tid == 0) {
                    const unsigned val = P.tag_base + (unsigned)a + 1u;
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + c * MAXCH3 + ch), "r"(val) : "memory");
                }
                if (k + 2 < P.Tseg && ch == 0) {   // gx rows of the step after next -> L2
                    const int tn = A.reverse ? t - 2 : t + 2;
                    const int nan = NA(tn);
                    const long long rbn = OFF(tn);
                    for (int n = gw; n < nan; n += GW) {
                        const float* gp = A.gx + (size_t)(rbn + (long long)n * ns + sl) * A.ld_gx + UN * c;
                        if (lane < 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + lane * HH));
                    }
                }
            }
            na_prev = na;
            ++a;
        }
    } else if (warp == GW) {
        // =========================== PRODUCER warp: flags + TMA ===========================
        const bool leader = elect_one();
        int q = 0, na_prev = 0, a = 0;
        for (int k = 0; k < P.Tseg; ++k) {
            const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
            const int na = NA(t);
            if (na == 0) {
                if (A.reverse) continue;
                break;
            }
            const int npoll = a == 0 ? 0 : min(na, na_prev);
            const int tp = A.reverse ? t + 1 : t - 1;                      // the step whose output rows are h_prev
            const long long prev_base = (tp >= 0 && tp < P.Ttot) ? OFF(tp) : 0;
            for (int ch = 0; ch * CN < na; ++ch, ++q) {
                const int slot = q % NS, u = q / NS;
                if (u >= 1) mbar_wait(dfull(slot), (uint32_t)((u - 1) & 1));   // the MMAs that read this slot last are complete
                if (ch * CN < npoll) {
                    const unsigned want = P.tag_base + (unsigned)a;            // written after active step a - 1
                    if (lane < CL) {
                        const unsigned* fp = flags + lane * MAXCH3 + ch;
                        const long long tp0 = clock64();
                        unsigned f;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(fp) : "memory");
                            POLL_GUARD(tp0);
                        } while ((int)(f - want) < 0 || (f >> 12) != (P.tag_base >> 12));
                    }
                    __syncwarp();
                }
                if (leader) {
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    mbar_expect_tx(full(slot), (uint32_t)(CN * 1024));
                    const long long row0 = prev_base + (long long)(ch * CN) * ns + sl;
                    const int y0 = (int)(row0 % ns), z0 = (int)(row0 / ns);
                    const uint32_t Hs = Hs0 + (uint32_t)(slot * CN * 1024);
#pragma unroll
                    for (int kb = 0; kb < 8; ++kb) tma_load_3d(Hs + (uint32_t)(kb * CN * 128), &tmH, XP.xcol[d] + 64 * kb, y0, z0, full(slot));
                }
                __syncwarp();
            }
            na_prev = na;
            ++a;
        }
    } else {
        // =========================== MMA warp ===========================
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc(CN);
        int q = 0;
        for (int k = 0; k < P.Tseg; ++k) {
            const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
            const int na = NA(t);
            if (na == 0) {
                if (A.reverse) continue;
                break;
            }
            for (int ch = 0; ch * CN < na; ++ch, ++q) {
                const int slot = q % NS, u = q / NS;
                const uint32_t Hs = Hs0 + (uint32_t)(slot * CN * 1024);
                const uint32_t td = tmem + TM_D + (uint32_t)(slot * CN);
                if (u >= 1) mbar_wait(dfree(slot), (uint32_t)((u - 1) & 1));
                mbar_wait(full(slot), (uint32_t)(u & 1));
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint64_t db = make_desc_k(Hs + (uint32_t)((j >> 2) * (CN * 128) + (j & 3) * 32));
                        tc_mma_ts(td, tmem + TM_A + 8 * j, db, idesc, j > 0 ? 1u : 0u);
                    }
                    tc_commit(dfull(slot));
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// =========================================================================================
// backward (BPTT).  Dual decomposition as in gru_mma.cu: dh_prev^T[512 x n] = R_own^T[512 x 96] . dgh_own^T[96 x n] from
// the CTA's OWN 96 dgh columns, the 512 partial sums reduce-scattered to their owner CTAs through LL words.  Here
// A = R_own^T sits in tensor memory as 4 M-tiles of 128 output units x 96 k (48 columns each), B = dgh_own (CN rows x
// 96 k, bf16) is written to swizzled shared memory by the gate-gradient threads, 24 tcgen05.mma (4 tiles x 6 k-steps,
// N = CN) produce D[512 x CN] in TMEM.  TMEM lane quadrant q of tile T holds exactly the 32 units owned by CTA 4T+q,
// so the send of a (tile, quadrant) is one tcgen05.ld + coalesced 256-byte LL stores per row pair.
// A row pair of the exchange is (n, n+8) within each 16 rows, so the warp that finishes rows w, w+8 reads each word once.
// CN = rows per MMA: 16, 32 or 64 (4 x CN accumulator columns + 192 operand columns <= 512).
// =========================================================================================
struct TcBwdDirP {
    const float* dhs; const float* hs_f; const bf16* hs_h; const float* h0; const float* cache; const bf16* R;
    float* dgx_f; bf16* dgx_h; float* dgh_f; bf16* dgh_h; float* hp_f; bf16* hp_h; float* dh0;
    const float* dh_in; float* dh_out;
    int ld_dhs, ld_hs, ld_dg, ld_hp, reverse;
};
struct TcBwdP {
    TcBwdDirP dir[2];
    const int* off; const int* nact;
    unsigned long long* ybuf;
    int ndir, nslices, b, Ttot, t0, Tseg, bslr;
    unsigned tag_base;
    long long* prof;
};
__device__ __forceinline__ uint2 ll_load1(const unsigned long long* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf16_bits(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16(x)); }
__device__ __forceinline__ size_t yidx(int par, int dest, int src, int pair, int ul, int npair) {
    return ((((size_t)par * CL + dest) * CL + src) * npair + pair) * UN + ul;
}
constexpr uint32_t TB_A = 0;      // columns [0,192): R_own^T, tile T at 48 T
constexpr uint32_t TB_D = 192;    // columns [192, 192 + 4 CN): accumulators, tile T at CN T

template <int CN>
__global__ void __launch_bounds__(NTH, 1) k_gru_tc_bwd(const __grid_constant__ TcBwdP P) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    if ((int)(blockIdx.x / CL) >= P.ndir * P.nslices) return;
    const uint32_t sbase = (smem_u32(sm_raw) + 1023u) & ~1023u;
    unsigned char* const sm = sm_raw + (sbase - smem_u32(sm_raw));
    // [Gs: 2 k-blocks x CN x 128 B][cs: bslr*32 f32][tables][bar][slot]
    const uint32_t Gs = sbase;
    float* cs = reinterpret_cast<float*>(sm + 2 * CN * 128);
    int* s_nact = reinterpret_cast<int*>(cs + (size_t)P.bslr * UN);
    int* s_off = s_nact + P.Tseg + 2;
    unsigned long long* barp = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(s_off + P.Tseg + 2) + 15) & ~uintptr_t(15));
    const uint32_t dbar = smem_u32(barp);
    const uint32_t tslot = dbar + 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = blockIdx.x / CL, c = blockIdx.x % CL;
    const int ns = P.nslices, d = grp / ns, sl = grp % ns;
    const TcBwdDirP& A = P.dir[d];
    const int npair = P.bslr / 2;
    for (int i = tid; i < P.Tseg + 2; i += NTH) {
        const int tt = P.t0 - 1 + i;
        s_nact[i] = (tt >= 0 && tt < P.Ttot) ? slice_rows(P.nact[tt], sl, ns) : 0;
        s_off[i] = (tt >= 0 && tt <= P.Ttot) ? P.off[tt] : 0;
    }
    if (tid == 0) {
        mbar_init(dbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");

    // ---- R_own^T -> tensor memory: tile T, lane 32q + l = output unit 128T + 32q + l, column j = local gate rows 2j, 2j+1
    if (warp < 4) {
        const unsigned short* Rs = reinterpret_cast<const unsigned short*>(A.R);
#pragma unroll 1
        for (int T = 0; T < 4; ++T) {
            const int o = 128 * T + 32 * warp + lane;
            const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + TB_A + 48 * T;
#pragma unroll 1
            for (int j8 = 0; j8 < 6; ++j8) {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lr = 2 * (8 * j8 + i);     // even local gate row; lr + 1 is the same gate
                    const size_t g0 = (size_t)((lr >> 5) * HH + UN * c + (lr & 31)) * HH + o;
                    r[i] = (uint32_t)Rs[g0] | ((uint32_t)Rs[g0 + HH] << 16);
                }
                tc_st8(tbase + 8 * j8, r);
            }
        }
        tc_wait_st();
    }
    const int col = UN * c + lane;
    const int nloc = slice_rows(P.b, sl, ns);
    for (int i = tid; i < nloc * UN; i += NTH)
        cs[i] = A.dh_in ? A.dh_in[(size_t)((i >> 5) * ns + sl) * HH + UN * c + (i & 31)] : 0.f;
    const size_t ypar = (size_t)CL * CL * npair * UN;
    unsigned long long* Y = P.ybuf + (size_t)grp * 2 * ypar;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t idesc = make_idesc(CN);
    constexpr int RPT = CN / 8;     // rows per thread: n = warp + 8 e
    int na_prev = 0;
    bool first = true;
    int k_last = -1;
    uint32_t dphase = 0;
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long plast = clock64();
    for (int k = 0; k < P.Tseg; ++k) {
        const int t = A.reverse ? P.t0 + k : P.t0 + P.Tseg - 1 - k;
        const int na = NA(t);
        if (na == 0) {
            if (A.reverse) break;
            continue;
        }
        PROF_MARK(0);
        const int ncarry = first ? 0 : min(na, na_prev);
        const int th = A.reverse ? t + 1 : t - 1;     // the step processed BEFORE t in the forward pass: source of h_prev
        int nhp = 0;
        long long hp_base = 0;
        bool hp_from_h0 = false;
        if (th >= 0 && th < P.Ttot) {
            nhp = min(na, NA(th));
            hp_base = OFF(th);
        } else if (!A.reverse && A.h0) {
            nhp = na;
            hp_from_h0 = true;
        }
        const unsigned tagr = P.tag_base + (unsigned)(k - 1), tagw = P.tag_base + (unsigned)k;
        const int parr = (k - 1) & 1, parw = k & 1;
        const long long row_base = OFF(t);
        for (int ch = 0; ch * CN < na; ++ch) {
            const int nrows = min(CN, na - ch * CN);
            // ---- this chunk's gate inputs (L2 hits: prefetched two steps ago), issued in front of the poll
            float in_dhs[RPT], in_r[RPT], in_z[RPT], in_n[RPT], in_q[RPT], in_hp[RPT];
#pragma unroll
            for (int e = 0; e < RPT; ++e) {
                const int n = warp + 8 * e;
                in_dhs[e] = in_r[e] = in_z[e] = in_n[e] = in_q[e] = in_hp[e] = 0.f;
                if (n < nrows) {
                    const int jl = ch * CN + n;
                    const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                    in_dhs[e] = ld_f32(A.dhs + row * A.ld_dhs + col);
                    const float* cp = A.cache + row * 4 * HH + col;
                    in_r[e] = ld_f32(cp); in_z[e] = ld_f32(cp + HH); in_n[e] = ld_f32(cp + 2 * HH); in_q[e] = ld_f32(cp + 3 * HH);
                    if (jl < nhp) {
                        if (hp_from_h0) {
                            in_hp[e] = ld_f32(A.h0 + (size_t)(jl * ns + sl) * HH + col);
                        } else {
                            const size_t rh = (size_t)(hp_base + (long long)jl * ns + sl);
                            in_hp[e] = A.hs_f ? ld_f32(A.hs_f + rh * A.ld_hs + col) : __bfloat162float(A.hs_h[rh * A.ld_hs + col]);
                        }
                    }
                }
            }
            // ---- reduce-scatter receive: partial sums of R^T.dgh for my unit, rows (w + 16 pe, w + 16 pe + 8), from all 16 CTAs
            float pin[RPT];
#pragma unroll
            for (int e = 0; e < RPT; ++e) pin[e] = 0.f;
#pragma unroll
            for (int pe = 0; pe < CN / 16; ++pe) {
                const int nlo = 16 * pe + warp;
                if (nlo < nrows && ch * CN + nlo < ncarry) {
                    const int pair = ch * (CN / 2) + 8 * pe + warp;
                    uint2 w[CL];
                    bool ok;
                    const long long tp0 = clock64();
                    const unsigned long long* yb = Y + yidx(parr, c, 0, pair, lane, npair);
                    const size_t ystr = (size_t)npair * UN;
                    do {
                        ok = true;
#pragma unroll
                        for (int s = 0; s < CL; ++s) w[s] = ll_load1(yb + s * ystr);
#pragma unroll
                        for (int s = 0; s < CL; ++s)
                            if (w[s].y != tagr) ok = false;
                        POLL_GUARD(tp0);
                    } while (!ok);
                    float lo = 0.f, hi = 0.f;
#pragma unroll
                    for (int s = 0; s < CL; ++s) { lo += bf16_lo(w[s].x); hi += bf16_hi(w[s].x); }
                    pin[2 * pe] = lo;
                    pin[2 * pe + 1] = hi;
                }
            }
            PROF_MARK(1);   // receive
            // ---- gate gradients of the rows this thread finishes; dgh_own -> swizzled shared memory (B operand)
            float o_dr[RPT], o_du[RPT], o_dn[RPT], o_dnr[RPT];
#pragma unroll
            for (int e = 0; e < RPT; ++e) {
                const int n = warp + 8 * e;
                float dr = 0.f, du = 0.f, dn = 0.f, dnr = 0.f;
                if (n < nrows) {
                    const int jl = ch * CN + n;
                    const float carry = cs[jl * UN + lane] + (jl < ncarry ? pin[e] : 0.f);
                    const float r = in_r[e], z = in_z[e], nn = in_n[e], qq = in_q[e], hp = in_hp[e];
                    const float dd = carry + in_dhs[e];
                    dn = dd * (1.f - z) * (1.f - nn * nn);
                    du = dd * (hp - nn) * z * (1.f - z);
                    dr = dn * qq * r * (1.f - r);
                    dnr = dn * r;
                    cs[jl * UN + lane] = dd * z;
                }
                o_dr[e] = dr; o_du[e] = du; o_dn[e] = dn; o_dnr[e] = dnr;
                // k = lane (dr), 32 + lane (du) in k-block 0; k = lane (dn.r) in k-block 1
                const uint32_t rowb = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128);
                const uint32_t sw = (uint32_t)(n & 7);
                const uint32_t a0 = Gs + rowb + ((((uint32_t)lane >> 3) ^ sw) << 4) + ((lane & 7) << 1);
                const uint32_t a1 = Gs + rowb + (((((uint32_t)lane >> 3) + 4) ^ sw) << 4) + ((lane & 7) << 1);
                const uint32_t a2 = Gs + CN * 128 + rowb + ((((uint32_t)lane >> 3) ^ sw) << 4) + ((lane & 7) << 1);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a0), "h"(__bfloat16_as_ushort(__float2bfloat16(dr))) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a1), "h"(__bfloat16_as_ushort(__float2bfloat16(du))) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a2), "h"(__bfloat16_as_ushort(__float2bfloat16(dnr))) : "memory");
            }
            PROF_MARK(2);   // gate gradients
            fence_async_smem();
            __syncthreads();
            PROF_MARK(3);   // barrier
            // ---- D_T[128 x CN] = R_own^T tile T [128 x 96] . dgh_own^T : 24 MMAs by one thread
            if (warp == NTH / 32 - 1) {
                if (elect_one()) {
                    tc_fence_after();
                    // k-step major: consecutive instructions add into different accumulators (the 4 M-tiles)
#pragma unroll
                    for (int j = 0; j < 6; ++j)
#pragma unroll
                        for (int T = 0; T < 4; ++T) {
                            const uint64_t db = make_desc_k(Gs + (uint32_t)((j >> 2) * (CN * 128) + (j & 3) * 32));
                            tc_mma_ts(tmem + TB_D + CN * T, tmem + TB_A + 48 * T + 8 * j, db, idesc, j > 0 ? 1u : 0u);
                        }
                    tc_commit(dbar);
                }
                __syncwarp();
            }
            mbar_wait(dbar, dphase);
            dphase ^= 1u;
            tc_fence_after();
            PROF_MARK(4);   // MMA
            // ---- reduce-scatter send: warp -> lane quadrant q = warp & 3 of tiles T = warp >> 2 and (warp >> 2) + 2; the
            // quadrant's 32 lanes are the units of CTA 4T + q.  Word = rows (n, n + 8) of 16 as two bf16 + tag
            {
                const int q = warp & 3;
                const int npairs_live = (nrows + 15) / 16 * 8;     // pairs (n, n+8) with n < nrows: all 8 of every started 16
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int T = (warp >> 2) + 2 * tt;
                    const int dest = 4 * T + q;
                    const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + TB_D + (uint32_t)(CN * T);
                    uint32_t r[CN];
#pragma unroll
                    for (int i = 0; i < CN / 8; ++i) tc_ld8(ta + 8 * i, r + 8 * i);
                    tc_wait_ld();
#pragma unroll
                    for (int p = 0; p < CN / 2; ++p) {
                        const int nlo = (p >> 3) * 16 + (p & 7);
                        if (p < npairs_live && nlo < nrows)
                            ll_store(Y + yidx(parw, dest, c, ch * (CN / 2) + p, lane, npair),
                                     bf16_bits(__uint_as_float(r[nlo])) | (bf16_bits(__uint_as_float(r[nlo + 8])) << 16), tagw);
                    }
                }
                tc_fence_before();
            }
            PROF_MARK(5);   // send
            // ---- bookkeeping, off the serial chain: this step's gate gradients -> HBM (operands of the batched wgrad / dgrad GEMMs)
#pragma unroll
            for (int e = 0; e < RPT; ++e) {
                const int n = warp + 8 * e;
                if (n >= nrows) continue;
                const int jl = ch * CN + n;
                const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                const size_t o = row * A.ld_dg + col;
                const float dr = o_dr[e], du = o_du[e], dn = o_dn[e], dnr = o_dnr[e], hp = in_hp[e];
                if (A.dgx_f) { A.dgx_f[o] = dr; A.dgx_f[o + HH] = du; A.dgx_f[o + 2 * HH] = dn; }
                if (A.dgx_h) { A.dgx_h[o] = __float2bfloat16(dr); A.dgx_h[o + HH] = __float2bfloat16(du); A.dgx_h[o + 2 * HH] = __float2bfloat16(dn); }
                if (A.dgh_f) { A.dgh_f[o] = dr; A.dgh_f[o + HH] = du; A.dgh_f[o + 2 * HH] = dnr; }
                if (A.dgh_h) { A.dgh_h[o] = __float2bfloat16(dr); A.dgh_h[o + HH] = __float2bfloat16(du); A.dgh_h[o + 2 * HH] = __float2bfloat16(dnr); }
                if (A.hp_f) A.hp_f[row * A.ld_hp + col] = hp;
                if (A.hp_h) A.hp_h[row * A.ld_hp + col] = __float2bfloat16(hp);
            }
            // gate inputs of the BPTT step after next -> L2
            if (ch == 0 && k + 2 < P.Tseg) {
                const int tn = A.reverse ? t + 2 : t - 2;
                const int nan = NA(tn);
                const long long rbn = OFF(tn);
                const int thn = A.reverse ? tn + 1 : tn - 1;
                const long long rbh = (thn >= 0 && thn < P.Ttot) ? OFF(thn) : -1;
                for (int n = warp; n < nan; n += NTH / 32) {
                    const size_t rown = (size_t)(rbn + (long long)n * ns + sl);
                    if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.cache + rown * 4 * HH + UN * c + lane * HH));
                    if (lane == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.dhs + rown * A.ld_dhs + UN * c));
                    if (lane == 5 && rbh >= 0 && A.hs_h && n < NA(thn))
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.hs_h + (size_t)(rbh + (long long)n * ns + sl) * A.ld_hs + UN * c));
                }
            }
            PROF_MARK(6);   // bookkeeping
        }
        na_prev = na;
        first = false;
        k_last = k;
        PROF_MARK(7);
    }
    if (P.prof && tid == 0)
        for (int i = 0; i < 8; ++i) P.prof[blockIdx.x * 8 + i] = pacc[i];
    // ---- gradient wrt the state before the segment's first forward-pass step (see gru_mma.cu)
    const bool accum_dh = !A.reverse && P.t0 == 0;
    float* const dh_dst = A.reverse ? ((P.t0 + P.Tseg < P.Ttot) ? A.dh_out : nullptr) : ((P.t0 == 0) ? A.dh0 : A.dh_out);
    if (dh_dst && k_last >= 0) {
        const unsigned tagr = P.tag_base + (unsigned)k_last;
        const int parr = k_last & 1;
        for (int ch = 0; ch * CN < na_prev; ++ch) {
#pragma unroll 1
            for (int pe = 0; pe < CN / 16; ++pe) {
                const int nlo = 16 * pe + warp;
                const int jlo = ch * CN + nlo;
                if (jlo >= na_prev) continue;
                const int pair = ch * (CN / 2) + 8 * pe + warp;
                uint2 w[CL];
                bool ok;
                const long long tp0 = clock64();
                do {
                    ok = true;
#pragma unroll
                    for (int s = 0; s < CL; ++s) w[s] = ll_load1(Y + yidx(parr, c, s, pair, lane, npair));
#pragma unroll
                    for (int s = 0; s < CL; ++s)
                        if (w[s].y != tagr) ok = false;
                    POLL_GUARD(tp0);
                } while (!ok);
                float lo = 0.f, hi = 0.f;
#pragma unroll
                for (int s = 0; s < CL; ++s) { lo += bf16_lo(w[s].x); hi += bf16_hi(w[s].x); }
                {
                    const size_t di = (size_t)(jlo * ns + sl) * HH + col;
                    const float val = cs[jlo * UN + lane] + lo;
                    dh_dst[di] = accum_dh ? dh_dst[di] + val : val;
                }
                if (jlo + 8 < na_prev) {
                    const size_t di = (size_t)((jlo + 8) * ns + sl) * HH + col;
                    const float val = cs[(jlo + 8) * UN + lane] + hi;
                    dh_dst[di] = accum_dh ? dh_dst[di] + val : val;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// =========================================================================================
// one GRU time step for hidden sizes whose recurrent matrix does not fit on chip (H != 512, e.g. BASELINE configs[4],
// H = 2048: R is 25 MB per direction-layer).  The generic path spent three device operations per step (memset, split-K
// GEMM h.R^T with a TMA reduce-add epilogue, gate kernel); this kernel is the step: CTA c owns hidden units
// [32c, 32c+32) = 96 rows of R, streams them and the bf16 state of up to 64 rows through a 4-stage TMA ring (A = R_own
// as three 32-row boxes per k-block, B = h_prev), accumulates D[128 x 64] in tensor memory (SS form, M = 128, N = 64) and
// finishes the cell in the epilogue exactly like the persistent kernels (TMEM lane quadrant = gate, gates of a unit meet
// in shared memory).  No inter-CTA exchange: the launch boundary is the exchange; the bf16 state ping-pongs between two
// buffers because the other CTAs still read the old one.  grid = (H/32, row chunks of 64, directions).
// =========================================================================================
struct StepDirP {
    const float* gx; const float* bR; float* state_f; bf16* state_h_out; float* hs_f; bf16* hs_h; float* cache;
    int ld_gx, ld_hs, na;
};
struct StepP {
    StepDirP dir[2];
    int H;
    int gx_fresh;   // 1: gx was written by the kernel right before this one (decode): read it after griddepcontrol.wait
};
constexpr int STEP_STAGES = 3;         // 3 x 24 KB + the gate tile: two CTAs fit an SM, so the NEXT step's CTAs (programmatic dependent
                                      // launch) become resident and prefetch their R blocks while this step still runs; 4 and 8 stages
                                      // measured the same step time (the stream is L2-bandwidth-, not latency-bound)
constexpr int STEP_A_BYTES = 128 * 128, STEP_B_BYTES = 64 * 128, STEP_STAGE_BYTES = STEP_A_BYTES + STEP_B_BYTES;
constexpr int STEP_NTH = 10 * 32;      // warp 0 producer, warp 1 MMA, warps 2..9 gate
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// MC = 8: clusters of 8 CTAs along the unit dimension share the state tile -- CTA r of a cluster loads rows [8r, 8r+8) of every
// k-block and the TMA engine multicasts them into all 8 CTAs (every CTA still streams its own R rows): 40 % less L2 traffic.
// A stage is free again when the MMA warps of ALL 8 CTAs have committed it (multicast commit onto everybody's empty barrier).
template <int MC>
__global__ void __launch_bounds__(STEP_NTH, 1) k_gru_step_fwd(const __grid_constant__ StepP P, const __grid_constant__ CUtensorMap tmR0,
                                                              const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmS0,
                                                              const __grid_constant__ CUtensorMap tmS1) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int c = blockIdx.x, chunk = blockIdx.y, d = blockIdx.z;
    const StepDirP& A = P.dir[d];
    const int r0 = chunk * 64;
    if (r0 >= A.na) return;                      // no live row in this chunk of this direction
    const int nrows = min(64, A.na - r0);
    const CUtensorMap* tmR = d ? &tmR1 : &tmR0;
    const CUtensorMap* tmS = d ? &tmS1 : &tmS0;
    const uint32_t sbase = (smem_u32(sm_raw) + 1023u) & ~1023u;
    unsigned char* const sm = sm_raw + (sbase - smem_u32(sm_raw));
    // [ring: 4 x (A 16 KB + B 8 KB)][G: 3*64*32 f32][barriers: full[4], empty[4], dfull][slot]
    float* G = reinterpret_cast<float*>(sm + STEP_STAGES * STEP_STAGE_BYTES);
    const uint32_t bars = sbase + STEP_STAGES * STEP_STAGE_BYTES + 3 * 64 * UN * 4;
    auto full = [&](int st) { return bars + 8u * (uint32_t)st; };
    auto empty = [&](int st) { return bars + 8u * (uint32_t)(STEP_STAGES + st); };
    const uint32_t dfull = bars + 8u * (2 * STEP_STAGES);
    const uint32_t tslot = dfull + 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = P.H, nkb = H / 64;
    if (tid == 0) {
        for (int i = 0; i < STEP_STAGES; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), MC); }
        mbar_init(dfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(tmR) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(tmS) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");
    uint32_t crank = 0;
    if (MC > 1) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        cluster_sync_all();     // every CTA's barriers are initialised before anybody multicasts into it
    }

    // Programmatic dependent launch: the next step's grid may start now (its prologue and its first R blocks do not depend on this
    // step); everything that does -- the state it reads, the state buffer it overwrites -- comes after griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 0) {
        // ---- producer: R_own (three gate blocks of 32 rows) and the state rows of the chunk, one k-block per stage
        const bool leader = elect_one();
        const int npre = min(nkb, STEP_STAGES);
        if (leader) {   // the first stages' R blocks: independent of the previous step
            for (int i = 0; i < npre; ++i) {
                const uint32_t sa = sbase + (uint32_t)(i * STEP_STAGE_BYTES);
                mbar_expect_tx(full(i), (uint32_t)(3 * 32 * 128 + 64 * 128));
#pragma unroll
                for (int g = 0; g < 3; ++g) tma_load_2d(sa + (uint32_t)(g * 32 * 128), tmR, 64 * i, g * H + UN * c, full(i));
            }
        }
        __syncwarp();
        asm volatile("griddepcontrol.wait;" ::: "memory");
        for (int i = 0; i < nkb; ++i) {
            const int st = i % STEP_STAGES;
            const uint32_t ph = (uint32_t)(i / STEP_STAGES) & 1u;
            if (i >= npre) mbar_wait(empty(st), ph ^ 1u);
            if (leader) {
                const uint32_t sa = sbase + (uint32_t)(st * STEP_STAGE_BYTES), sb = sa + STEP_A_BYTES;
                if (i >= npre) {
                    mbar_expect_tx(full(st), (uint32_t)(3 * 32 * 128 + 64 * 128));
#pragma unroll
                    for (int g = 0; g < 3; ++g) tma_load_2d(sa + (uint32_t)(g * 32 * 128), tmR, 64 * i, g * H + UN * c, full(st));
                }
                if (MC > 1) tma_load_2d_mc(sb + crank * 1024u, tmS, 64 * i, r0 + 8 * (int)crank, full(st), (uint16_t)((1u << MC) - 1u));
                else tma_load_2d(sb, tmS, 64 * i, r0, full(st));
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---- MMA: D[128 x 64] += A[128 x 16] . B[64 x 16]^T, four per k-block (rows 96..127 of A are never loaded: their
        // accumulator lanes are garbage nobody reads)
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc(64);
        for (int i = 0; i < nkb; ++i) {
            const int st = i % STEP_STAGES;
            const uint32_t ph = (uint32_t)(i / STEP_STAGES) & 1u;
            mbar_wait(full(st), ph);
            tc_fence_after();
            if (leader) {
                const uint32_t sa = sbase + (uint32_t)(st * STEP_STAGE_BYTES), sb = sa + STEP_A_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) tc_mma_ss(tmem, make_desc_k(sa + j * 32), make_desc_k(sb + j * 32), idesc, (i > 0 || j > 0) ? 1u : 0u);
                if (MC > 1) tc_commit_mc(empty(st), (uint16_t)((1u << MC) - 1u));
                else tc_commit(empty(st));
                if (i == nkb - 1) tc_commit(dfull);
            }
            __syncwarp();
        }
    } else {
        // ---- gate warps: gw = 0..7, lane quadrant = gate, two column halves
        const int gw = warp - 2, quad = warp & 3, half = gw >> 2;
        const int col = UN * c + lane;
        const float bRr = A.bR[col], bRu = A.bR[H + col], bRn = A.bR[2 * H + col];
        float gxv[8][3], hpv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int n = gw + 8 * e;
            gxv[e][0] = gxv[e][1] = gxv[e][2] = 0.f; hpv[e] = 0.f;
            if (n < nrows && !P.gx_fresh) {
                const float* gp = A.gx + (size_t)(r0 + n) * A.ld_gx + col;
                gxv[e][0] = ld_f32(gp); gxv[e][1] = ld_f32(gp + H); gxv[e][2] = ld_f32(gp + 2 * H);
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");     // the previous step's state is complete (and nobody reads its input any more)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int n = gw + 8 * e;
            if (n < nrows) {
                hpv[e] = ld_f32(A.state_f + (size_t)(r0 + n) * H + col);
                if (P.gx_fresh) {
                    const float* gp = A.gx + (size_t)(r0 + n) * A.ld_gx + col;
                    gxv[e][0] = ld_f32(gp); gxv[e][1] = ld_f32(gp + H); gxv[e][2] = ld_f32(gp + 2 * H);
                }
            }
        }
        if (quad < 3) {
            mbar_wait(dfull, 0);
            tc_fence_after();
            const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32);
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 4; ++i) tc_ld8(ta + 8 * i, r + 8 * i);
            tc_wait_ld();
            tc_fence_before();
            float* gp = G + ((size_t)quad * 64 + half * 32) * UN + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) gp[i * UN] = __uint_as_float(r[i]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int n = gw + 8 * e;
            if (n < nrows) {
                const float s0 = G[(0 * 64 + n) * UN + lane], s1 = G[(1 * 64 + n) * UN + lane], s2 = G[(2 * 64 + n) * UN + lane];
                // same arithmetic as the generic gate kernel (expf / tanhf): the validation paths compare against it
                const float r = 1.f / (1.f + expf(-(gxv[e][0] + s0 + bRr)));
                const float z = 1.f / (1.f + expf(-(gxv[e][1] + s1 + bRu)));
                const float qq = s2 + bRn;
                const float nv = tanhf(gxv[e][2] + r * qq);
                const float h = (1.f - z) * nv + z * hpv[e];
                const size_t row = (size_t)(r0 + n);
                A.state_f[row * H + col] = h;
                const __nv_bfloat16 hb = __float2bfloat16(h);
                A.state_h_out[row * H + col] = hb;
                if (A.hs_h) A.hs_h[row * A.ld_hs + col] = hb;
                if (A.hs_f) A.hs_f[row * A.ld_hs + col] = h;
                if (A.cache) {
                    float* cp = A.cache + row * 4 * H + col;
                    cp[0] = r; cp[H] = z; cp[2 * H] = nv; cp[3 * H] = qq;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MC > 1) cluster_sync_all();     // nobody leaves while a peer may still arrive on its barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
}

// =========================================================================================
// unit test of the TS-form building blocks: D[128 x N] = A[128 x K] . B[N x K]^T with A written to tensor memory by
// tcgen05.st, B staged in swizzled shared memory by plain stores, K = 64 * kblocks
// =========================================================================================
template <int NACC, int NMMA>
__device__ __forceinline__ void issue_unrolled(uint32_t tmem, uint32_t Hs, int N, uint32_t idesc) {
#pragma unroll
    for (int j = 0; j < NMMA; ++j) {
        const uint64_t db = make_desc_k(Hs + (uint32_t)((j >> 2) * (N * 128) + (j & 3) * 32));
        tc_mma_ts(tmem + TM_D + (uint32_t)((j % NACC) * N), tmem + TM_A + 8 * j, db, idesc, j >= NACC ? 1u : 0u);
    }
}
__global__ void __launch_bounds__(128, 1) k_test_ts_mma(const bf16* __restrict__ Ag, const bf16* __restrict__ Bg, float* __restrict__ Dg,
                                                        int N, int K, int nacc, long long* __restrict__ cycles) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const uint32_t sbase = (smem_u32(sm_raw) + 1023u) & ~1023u;
    const uint32_t Hs = sbase;
    const uint32_t dbar = sbase + (uint32_t)(N * K * 2);
    const uint32_t tslot = dbar + 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(dbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");
    {   // A: thread = row (lane of tensor memory), K/2 columns
        const uint4* src = reinterpret_cast<const uint4*>(Ag + (size_t)tid * K);
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + TM_A;
        for (int j = 0; j < K / 16; ++j) {
            const uint4 v0 = src[2 * j], v1 = src[2 * j + 1];
            uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            tc_st8(tbase + 8 * j, r);
        }
        tc_wait_st();
    }
    // B: (N x K) row-major bf16 -> [k-block][N rows x 128 B], 128-byte swizzle
    for (int idx = tid; idx < N * (K / 8); idx += 128) {
        const int n = idx / (K / 8), cidx = idx % (K / 8);
        const uint4 v = *reinterpret_cast<const uint4*>(Bg + (size_t)n * K + 8 * cidx);
        const uint32_t off = (uint32_t)((cidx >> 3) * (N * 128) + (n >> 3) * 1024 + (n & 7) * 128 + (((cidx & 7) ^ (n & 7)) << 4));
        st_shared_v4(Hs + off, v.x, v.y, v.z, v.w);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // the K/16 MMAs go round robin over `nacc` accumulators (column ranges of N): consecutive instructions that add
    // into the SAME accumulator wait for each other, independent ones pipeline
    const long long c0k = clock64();
    if (warp == 0 && elect_one()) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(N);
        if (K == 512 && nacc == 1) issue_unrolled<1, 32>(tmem, Hs, N, idesc);        // the forms the recurrence kernels use:
        else if (K == 512 && nacc == 2) issue_unrolled<2, 32>(tmem, Hs, N, idesc);   // fully unrolled, constant operands
        else if (K == 512 && nacc == 4) issue_unrolled<4, 32>(tmem, Hs, N, idesc);
        else if (K == 512 && nacc == 8) issue_unrolled<8, 32>(tmem, Hs, N, idesc);
        else if (K == 512 && nacc == 16) issue_unrolled<16, 32>(tmem, Hs, N, idesc);
        else
            for (int j = 0; j < K / 16; ++j) {
                const uint64_t db = make_desc_k(Hs + (uint32_t)((j >> 2) * (N * 128) + (j & 3) * 32));
                tc_mma_ts(tmem + TM_D + (uint32_t)((j % nacc) * N), tmem + TM_A + 8 * j, db, idesc, j >= nacc ? 1u : 0u);
            }
        if (cycles) cycles[1] = clock64() - c0k;     // issue time
        tc_commit(dbar);
    }
    __syncwarp();
    mbar_wait(dbar, 0);
    tc_fence_after();
    if (tid == 0 && cycles) cycles[0] = clock64() - c0k;
    for (int c0 = 0; c0 < N; c0 += 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int a = 0; a < nacc && a < K / 16; ++a) {
            uint32_t r[8];
            tc_ld8(tmem + ((uint32_t)(warp * 32) << 16) + TM_D + (uint32_t)(a * N) + c0, r);
            tc_wait_ld();
            for (int i = 0; i < 8; ++i) acc[i] += __uint_as_float(r[i]);
        }
        for (int i = 0; i < 8; ++i) Dg[(size_t)tid * N + c0 + i] = acc[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <int CN>
size_t fwd2_smem(int bslr, int Tseg) {
    return 1024 + (size_t)2 * CN * 1024 + (size_t)2 * 3 * CN * UN * 4 + (size_t)bslr * UN * 4 + (size_t)(2 * Tseg + 4) * 4 + 16 + 14 * 8 + 16;
}
template <int CN, int NS>
size_t fwd3_smem(int bslr, int Tseg) {
    return 1024 + (size_t)NS * CN * 1024 + (size_t)3 * CN * UN * 4 + (size_t)bslr * UN * 4 + (size_t)(2 * Tseg + 4) * 4 + 16 + (3 * NS + 2) * 8 + 16;
}
template <int CN>
size_t bwd_smem(int bslr, int Tseg) {
    return 1024 + (size_t)2 * CN * 128 + (size_t)bslr * UN * 4 + (size_t)(2 * Tseg + 4) * 4 + 16 + 16 + 16;
}
template <int CN>
size_t fwd_smem(int bslr, int Tseg) {
    return 1024 + (size_t)CN * 1024 + (size_t)3 * CN * UN * 4 + (size_t)bslr * UN * 4 + (size_t)(2 * Tseg + 4) * 4 + 16 + 16 + 16;
}

}  // namespace

struct GruTcCtx {
    int device = 0, num_sms = 148;
    unsigned launch_id = 1;
    static constexpr int NSLOT = 4;
    unsigned long long* xbuf[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    size_t xcap[NSLOT] = {0, 0, 0, 0};
    unsigned long long* ybuf[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    size_t ycap[NSLOT] = {0, 0, 0, 0};
    unsigned* flags[NSLOT] = {nullptr, nullptr, nullptr, nullptr};   // forward kernel 3: [group][CL][MAXCH3] step flags
    int fwd3_cn = 6416;          // ARGSIM_GRU_TC_FWD3_CN: variant of the TMA-fed kernel.  Recurrence of the first embedding micro-batch (1,024
                                 // rows, 153 steps, 3 layers): 6416 = two slots of 64 rows, 16 gate warps: 2.75 ms; 48 = three slots of 48 rows, 16
                                 // gate warps: 3.02 ms; 64 = two slots of 64 rows, 8 gate warps: 3.09-3.22 ms; 32 = four slots of 32 rows, 8 gate
                                 // warps: 3.56 ms (the per-chunk costs -- flag round trip, 32 MMA issues -- do not shrink with the chunk; the
                                 // gate math does shrink with the number of gate warps)
    int fwd3_min_rows = 100;      // ARGSIM_GRU_TC_FWD3_ROWS: rows per slice from which whole-layer launches take the TMA-fed kernel
    int pad_groups = 8;
    int force_cn = 0;            // ARGSIM_GRU_TC_CN: rows per MMA chunk (16 / 32 / 64 / 128), 0 = by live rows
    int fwd_version = 1;         // LL-exchange forward kernel for launches the TMA-fed kernel does not take: 1 = block-synchronous
                                 // k_gru_tc_fwd (2,700 cycles per C1 step), ARGSIM_GRU_TC_FWD=2 = warp-specialised k_gru_tc_fwd2 (3,230)
    bool fwd3_on = true;         // ARGSIM_GRU_TC_FWD=0: no TMA-fed throughput kernel
    int poll_delay = 300;        // ARGSIM_GRU_TC_DELAY: cycles between the local publish and a stage warp's first poll round
    long long* prof = nullptr;   // ARGSIM_GRU_PROF=1
};

GruTcCtx* gru_tc_create(int device) {
    GruTcCtx* c = new GruTcCtx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    const int max_smem = 232448;   // 227 KB: the launcher checks the actual request against it
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd2<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd2<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    if (const char* v = getenv("ARGSIM_GRU_TC_FWD")) {
        c->fwd_version = atoi(v) == 2 ? 2 : 1;
        c->fwd3_on = atoi(v) != 0;
    }
    if (const char* v = getenv("ARGSIM_GRU_TC_FWD3_ROWS")) c->fwd3_min_rows = atoi(v);
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd3<64, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd3<32, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd3<48, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_fwd3<64, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    if (const char* v = getenv("ARGSIM_GRU_TC_FWD3_CN")) c->fwd3_cn = atoi(v);
    if (const char* v = getenv("ARGSIM_GRU_TC_DELAY")) c->poll_delay = atoi(v);
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_bwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_bwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_tc_bwd<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    if (const char* v = getenv("ARGSIM_GRU_PAD")) c->pad_groups = atoi(v);
    if (const char* v = getenv("ARGSIM_GRU_TC_CN")) c->force_cn = atoi(v);
    if (getenv("ARGSIM_GRU_PROF")) {
        CUDA_CHECK(cudaMalloc(&c->prof, 160 * 16 * sizeof(long long)));
        CUDA_CHECK(cudaMemset(c->prof, 0, 160 * 16 * sizeof(long long)));
    }
    return c;
}
void gru_tc_destroy(GruTcCtx* c) {
    if (!c) return;
    for (int i = 0; i < GruTcCtx::NSLOT; ++i) { cudaFree(c->xbuf[i]); cudaFree(c->ybuf[i]); cudaFree(c->flags[i]); }
    cudaFree(c->prof);
    delete c;
}
bool gru_tc_supported(int H) { return H == HH; }

// slices and rows per MMA chunk for `b` live rows of `ndir` directions: as many groups as fit (9 groups of 16 CTAs),
// at least 16 rows per slice (N = 16 is the smallest M = 128 instruction); rows_per_slice > 0 forces the slice size
static void tc_pick(const GruTcCtx* c, int ndir, int b, int rows_per_slice, int* ns, int* bslr, int* cn) {
    const int max_groups = std::max(1, c->num_sms / CL);
    const int per_want = rows_per_slice > 0 ? rows_per_slice : 16;
    int s = std::max(1, std::min(max_groups / ndir, (b + per_want - 1) / per_want));
    const int per = (b + s - 1) / s;
    int n = per <= 16 ? 16 : per <= 32 ? 32 : per <= 64 ? 64 : 128;
    if (c->force_cn == 16 || c->force_cn == 32 || c->force_cn == 64 || c->force_cn == 128) n = std::min(n, c->force_cn);
    *ns = s;
    *cn = n;
    *bslr = (per + n - 1) / n * n;
}
// true when a whole-layer launch of `b` rows would take the TMA-fed throughput kernel (rows per slice >= fwd3_min_rows)
bool gru_tc_throughput(const GruTcCtx* c, int ndir, int b) {
    int ns, bslr, cn;
    tc_pick(c, ndir, b, 0, &ns, &bslr, &cn);
    const int per = (b + ns - 1) / ns;
    return c->fwd3_on && per >= c->fwd3_min_rows && per <= 256;
}
bool gru_tc_fits(const GruTcCtx* c, int ndir, int b) {
    int ns, bslr, cn;
    tc_pick(c, ndir, b, 0, &ns, &bslr, &cn);
    return ndir * CL <= c->num_sms && bslr <= MAX_BSL;
}

void gru_tc_fwd(GruTcCtx* c, const GruFwdArgs* dirs, int ndir, const SeqPlan& Pl, const int* d_off, const int* d_nact, int H,
                cudaStream_t s, int t0, int Tseg, int slot, int rows_per_slice, int pad) {
    if (H != HH) throw std::runtime_error("gru_tc: H must be 512");
    if (Tseg < 0) { t0 = 0; Tseg = Pl.Tmax; }
    if (Tseg >= 4096) throw std::runtime_error("gru_tc: more than 4095 steps per launch");
    if (slot < 0 || slot >= GruTcCtx::NSLOT) throw std::runtime_error("gru_tc: bad slot");
    TcFwdP P;
    int ns, bslr, cn;
    const int b_seg = (ndir == 1) ? Pl.nact[t0] : Pl.b;
    tc_pick(c, ndir, b_seg, rows_per_slice, &ns, &bslr, &cn);
    if (c->fwd_version == 2 && cn > 64) {   // two operand slots of CN KB each
        cn = 64;
        const int per = (b_seg + ns - 1) / ns;
        bslr = (per + 63) / 64 * 64;
    }
    if (bslr > MAX_BSL) throw std::runtime_error("gru_tc: batch too large for the persistent kernel");
    for (int d = 0; d < ndir; ++d) {
        const GruFwdArgs& a = dirs[d];
        if (!a.R_h) throw std::runtime_error("gru_tc: bf16 weights missing");
        if (a.h0 && a.reverse && ndir != 1) throw std::runtime_error("gru_tc: a reverse direction takes h0 only in a single-direction (segment) launch");
        P.dir[d] = TcDirP{a.gx, a.R_h, a.bR, a.h0, a.hs_f, a.hs_h, a.cache, a.hT, a.ld_gx, a.ld_hs, a.reverse};
    }
    if (ndir == 1) P.dir[1] = P.dir[0];
    const int groups = ndir * ns;
    {   // throughput form: whole-layer launches with many rows per slice read h_{t-1} from the layer's output through TMA
        bool ok3 = c->fwd3_on && rows_per_slice == 0 && (b_seg + ns - 1) / ns >= c->fwd3_min_rows && t0 == 0 && Tseg == Pl.Tmax;
        for (int d = 0; d < ndir; ++d) ok3 = ok3 && dirs[d].hs_h && !dirs[d].h0 && !dirs[d].hT && dirs[d].ld_hs == dirs[0].ld_hs;
        long long xc1 = 0;
        if (ok3 && ndir == 2) {
            xc1 = dirs[1].hs_h - dirs[0].hs_h;
            ok3 = xc1 >= 0 && xc1 + HH <= dirs[0].ld_hs;
        }
        if (ok3) {
            const int per = (b_seg + ns - 1) / ns;
            // variant: rows per chunk x slots x gate warps.  48: 3 slots of 48 rows, 16 gate warps; 64 / 32: two / four slots, 8 gate warps;
            // 6416: 2 slots of 64 rows, 16 gate warps
            const int var = c->fwd3_cn;
            const int cn3 = var == 48 ? 48 : var == 32 ? 32 : 64;
            const int bslr3 = (per + cn3 - 1) / cn3 * cn3;
            if (bslr3 > MAX_BSL) throw std::runtime_error("gru_tc: batch too large for the persistent kernel");
            if (!c->flags[slot]) {
                CUDA_CHECK(cudaMalloc(&c->flags[slot], (size_t)(c->num_sms / CL + 1) * CL * MAXCH3 * sizeof(unsigned)));
                CUDA_CHECK(cudaMemset(c->flags[slot], 0, (size_t)(c->num_sms / CL + 1) * CL * MAXCH3 * sizeof(unsigned)));
                CUDA_CHECK(cudaDeviceSynchronize());
            }
            P.off = d_off; P.nact = d_nact; P.xbuf = nullptr;
            P.ndir = ndir; P.nslices = ns; P.b = b_seg; P.Ttot = Pl.Tmax; P.t0 = 0; P.Tseg = Tseg; P.bslr = bslr3;
            P.tag_base = (c->launch_id++) << 12;
            if (c->launch_id >= (1u << 20)) c->launch_id = 1;
            P.prof = nullptr; P.poll_delay = 0;
            CUtensorMap tm;
            tma_encode_slice_rows_bf16(&tm, dirs[0].hs_h, dirs[0].ld_hs, Pl.rows, ns, cn3);
            TcFwd3X X3;
            X3.flags = c->flags[slot]; X3.xcol[0] = 0; X3.xcol[1] = (int)xc1;
            void* args3[] = {&P, &tm, &X3};
            void* fn3; size_t smem3; int nth3;
            if (var == 48) { fn3 = (void*)k_gru_tc_fwd3<48, 3, 16>; smem3 = fwd3_smem<48, 3>(bslr3, Tseg); nth3 = 18 * 32; }
            else if (var == 32) { fn3 = (void*)k_gru_tc_fwd3<32, 4, 8>; smem3 = fwd3_smem<32, 4>(bslr3, Tseg); nth3 = 10 * 32; }
            else if (var == 6416) { fn3 = (void*)k_gru_tc_fwd3<64, 2, 16>; smem3 = fwd3_smem<64, 2>(bslr3, Tseg); nth3 = 18 * 32; }
            else { fn3 = (void*)k_gru_tc_fwd3<64, 2, 8>; smem3 = fwd3_smem<64, 2>(bslr3, Tseg); nth3 = 10 * 32; }
            if (smem3 > 232448) throw std::runtime_error("gru_tc: shared memory request exceeds 227 KB");
            const int gg = pad ? std::max(groups, c->pad_groups) : groups;
            if (pad == 2) CUDA_CHECK(cudaLaunchKernel(fn3, dim3(gg * CL), dim3(nth3), args3, smem3, s));
            else CUDA_CHECK(cudaLaunchCooperativeKernel(fn3, dim3(gg * CL), dim3(nth3), args3, smem3, s));
            COUNT_LAUNCH();
            return;
        }
    }
    const size_t need = (size_t)groups * 2 * bslr * (HH / 2);
    if (need > c->xcap[slot]) {
        CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(c->xbuf[slot]);
        CUDA_CHECK(cudaMalloc(&c->xbuf[slot], need * 8));
        CUDA_CHECK(cudaMemset(c->xbuf[slot], 0, need * 8));
        CUDA_CHECK(cudaDeviceSynchronize());
        c->xcap[slot] = need;
    }
    P.off = d_off; P.nact = d_nact; P.xbuf = c->xbuf[slot];
    P.ndir = ndir; P.nslices = ns; P.b = b_seg; P.Ttot = Pl.Tmax; P.t0 = t0; P.Tseg = Tseg; P.bslr = bslr;
    P.tag_base = (c->launch_id++) << 12;
    if (c->launch_id >= (1u << 20)) c->launch_id = 1;
    P.prof = c->prof;
    P.poll_delay = c->poll_delay;
    void* args[] = {&P};
    const int grid_groups = pad ? std::max(groups, c->pad_groups) : groups;
    void* fn; size_t smem;
    int nth = NTH;
    if (c->fwd_version == 2) {
        nth = NTH2;
        switch (cn) {
            case 16: fn = (void*)k_gru_tc_fwd2<16>; smem = fwd2_smem<16>(bslr, Tseg); break;
            case 32: fn = (void*)k_gru_tc_fwd2<32>; smem = fwd2_smem<32>(bslr, Tseg); break;
            default: fn = (void*)k_gru_tc_fwd2<64>; smem = fwd2_smem<64>(bslr, Tseg); break;
        }
    } else {
        switch (cn) {
            case 16: fn = (void*)k_gru_tc_fwd<16>; smem = fwd_smem<16>(bslr, Tseg); break;
            case 32: fn = (void*)k_gru_tc_fwd<32>; smem = fwd_smem<32>(bslr, Tseg); break;
            case 64: fn = (void*)k_gru_tc_fwd<64>; smem = fwd_smem<64>(bslr, Tseg); break;
            default: fn = (void*)k_gru_tc_fwd<128>; smem = fwd_smem<128>(bslr, Tseg); break;
        }
    }
    if (smem > 232448) throw std::runtime_error("gru_tc: shared memory request exceeds 227 KB (segment too long for this slice size)");
    if (pad == 2) CUDA_CHECK(cudaLaunchKernel(fn, dim3(grid_groups * CL), dim3(nth), args, smem, s));
    else CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(grid_groups * CL), dim3(nth), args, smem, s));
    COUNT_LAUNCH();
    if (c->prof) {
        CUDA_CHECK(cudaStreamSynchronize(s));
        const int W = c->fwd_version == 2 ? 16 : 8;
        std::vector<long long> h((size_t)groups * CL * W);
        CUDA_CHECK(cudaMemcpy(h.data(), c->prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg[16] = {0};
        for (int b2 = 0; b2 < groups * CL; ++b2)
            for (int i = 0; i < W; ++i) avg[i] += (double)h[b2 * W + i] / (groups * CL);
        fprintf(stderr, "[gru_tc_prof] fwd v%d ndir=%d ns=%d cn=%d steps=%d cycles/step:", c->fwd_version, ndir, ns, cn, Tseg);
        for (int i = 0; i < W; ++i) fprintf(stderr, " p%d=%.0f", i, avg[i] / Tseg);
        fprintf(stderr, "\n");
    }
}

void gru_tc_bwd(GruTcCtx* c, const GruBwdArgs* dirs, int ndir, const SeqPlan& Pl, const int* d_off, const int* d_nact, int H,
                cudaStream_t s, int t0, int Tseg, int slot, int rows_per_slice, int pad) {
    if (H != HH) throw std::runtime_error("gru_tc: H must be 512");
    if (Tseg < 0) { t0 = 0; Tseg = Pl.Tmax; }
    if (Tseg >= 4096) throw std::runtime_error("gru_tc: more than 4095 steps per launch");
    if (slot < 0 || slot >= GruTcCtx::NSLOT) throw std::runtime_error("gru_tc: bad slot");
    TcBwdP P;
    int ns, bslr, cn;
    const int b_seg = (ndir == 1) ? Pl.nact[t0] : Pl.b;
    tc_pick(c, ndir, b_seg, rows_per_slice, &ns, &bslr, &cn);
    if (cn > 64) {   // the accumulators of 4 M-tiles need 4 x CN columns next to the 192 operand columns
        cn = 64;
        const int per = (b_seg + ns - 1) / ns;
        bslr = (per + 63) / 64 * 64;
    }
    if (bslr > MAX_BSL) throw std::runtime_error("gru_tc: batch too large for the persistent kernel");
    for (int d = 0; d < ndir; ++d) {
        const GruBwdArgs& a = dirs[d];
        if (!a.R_h) throw std::runtime_error("gru_tc: bf16 weights missing");
        P.dir[d] = TcBwdDirP{a.dhs, a.hs_f, a.hs_h, a.h0, a.cache, a.R_h, a.dgx_f, a.dgx_h, a.dgh_f, a.dgh_h, a.hp_f, a.hp_h, a.dh0,
                             a.dh_in, a.dh_out, a.ld_dhs, a.ld_hs, a.ld_dg, a.ld_hp, a.reverse};
    }
    if (ndir == 1) P.dir[1] = P.dir[0];
    const int groups = ndir * ns;
    const size_t need = (size_t)groups * 2 * CL * CL * (bslr / 2) * UN;
    if (need > c->ycap[slot]) {
        CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(c->ybuf[slot]);
        CUDA_CHECK(cudaMalloc(&c->ybuf[slot], need * 8));
        CUDA_CHECK(cudaMemset(c->ybuf[slot], 0, need * 8));
        CUDA_CHECK(cudaDeviceSynchronize());
        c->ycap[slot] = need;
    }
    P.off = d_off; P.nact = d_nact; P.ybuf = c->ybuf[slot];
    P.ndir = ndir; P.nslices = ns; P.b = b_seg; P.Ttot = Pl.Tmax; P.t0 = t0; P.Tseg = Tseg; P.bslr = bslr;
    P.tag_base = (c->launch_id++) << 12;
    if (c->launch_id >= (1u << 20)) c->launch_id = 1;
    P.prof = c->prof;
    void* args[] = {&P};
    const int grid_groups = pad ? std::max(groups, c->pad_groups) : groups;
    void* fn; size_t smem;
    switch (cn) {
        case 16: fn = (void*)k_gru_tc_bwd<16>; smem = bwd_smem<16>(bslr, Tseg); break;
        case 32: fn = (void*)k_gru_tc_bwd<32>; smem = bwd_smem<32>(bslr, Tseg); break;
        default: fn = (void*)k_gru_tc_bwd<64>; smem = bwd_smem<64>(bslr, Tseg); break;
    }
    if (smem > 232448) throw std::runtime_error("gru_tc: shared memory request exceeds 227 KB");
    if (pad == 2) CUDA_CHECK(cudaLaunchKernel(fn, dim3(grid_groups * CL), dim3(NTH), args, smem, s));
    else CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(grid_groups * CL), dim3(NTH), args, smem, s));
    COUNT_LAUNCH();
    if (c->prof) {
        CUDA_CHECK(cudaStreamSynchronize(s));
        std::vector<long long> h((size_t)groups * CL * 8);
        CUDA_CHECK(cudaMemcpy(h.data(), c->prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg[8] = {0};
        for (int b2 = 0; b2 < groups * CL; ++b2)
            for (int i = 0; i < 8; ++i) avg[i] += (double)h[b2 * 8 + i] / (groups * CL);
        fprintf(stderr, "[gru_tc_prof] bwd ndir=%d ns=%d cn=%d steps=%d cycles/step:", ndir, ns, cn, Tseg);
        for (int i = 0; i < 8; ++i) fprintf(stderr, " p%d=%.0f", i, avg[i] / Tseg);
        fprintf(stderr, "\n");
    }
}

// ---- per-step fused GRU forward for hidden sizes without a persistent kernel (see k_gru_step_fwd) -------------------------
void tma_encode_2d_bf16(void* map_out, const bf16* base, int ld, long long rows, int cols, int box_rows);
bool gru_step_supported(int H, int b) { return H % 64 == 0 && H >= 64 && b >= 1; }
// state_f: (b,H) fp32 per direction (in/out); state_h[2]: (b,H) bf16 ping-pong per direction (state_h[cur] holds h_prev)
void gru_step_fwd(const GruFwdArgs* dirs, int ndir, const int* na, const long long* row0, int b, int H, float* const* state_f,
                  bf16* const* state_h_cur, bf16* const* state_h_next, cudaStream_t s, bool gx_fresh) {
    static bool configured = false;
    // measured on the scaled config: multicast 146.7 ms per step, plain 144.0 (cluster syncs and lock-step stages cost more than
    // the 40 % of L2 bytes they save: the stream is bound by every CTA's own R rows) -> off unless ARGSIM_STEP_MULTICAST=1
    static const bool no_mc = getenv("ARGSIM_STEP_MULTICAST") == nullptr;
    const size_t smem = 1024 + (size_t)STEP_STAGES * STEP_STAGE_BYTES + 3 * 64 * UN * 4 + (2 * STEP_STAGES + 1) * 8 + 16;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(k_gru_step_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_gru_step_fwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const bool mc = !no_mc && (H / UN) % 8 == 0;
    StepP P;
    P.H = H;
    P.gx_fresh = gx_fresh ? 1 : 0;
    CUtensorMap tmR[2], tmS[2];
    int na_max = 0;
    for (int d = 0; d < 2; ++d) {
        const int dd = d < ndir ? d : 0;
        const GruFwdArgs& a = dirs[dd];
        const long long r0 = row0[dd];
        P.dir[d] = StepDirP{a.gx + r0 * a.ld_gx, a.bR, state_f[dd], state_h_next[dd], a.hs_f ? a.hs_f + r0 * a.ld_hs : nullptr,
                            a.hs_h ? a.hs_h + r0 * a.ld_hs : nullptr, a.cache ? a.cache + r0 * 4 * H : nullptr, a.ld_gx, a.ld_hs,
                            d < ndir ? na[dd] : 0};
        tma_encode_2d_bf16(&tmR[d], a.R_h, H, 3LL * H, H, 32);
        tma_encode_2d_bf16(&tmS[d], state_h_cur[dd], H, b, H, mc ? 8 : 64);
        if (d < ndir) na_max = std::max(na_max, na[dd]);
    }
    if (na_max <= 0) return;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(H / UN, (na_max + 63) / 64, ndir);
    cfg.blockDim = dim3(STEP_NTH);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 8; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = mc ? 2 : 1;
    static const bool no_pdl = getenv("ARGSIM_STEP_NO_PDL") != nullptr;
    // gx written by the kernel right before (decode): nothing to overlap with, and the launch may sit in a captured graph
    if ((no_pdl || gx_fresh) && !mc) cfg.numAttrs = 0;
    if (mc) CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gru_step_fwd<8>, P, tmR[0], tmR[1], tmS[0], tmS[1]));
    else CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gru_step_fwd<1>, P, tmR[0], tmR[1], tmS[0], tmS[1]));
    COUNT_LAUNCH();
}

// D(128,N) = A(128,K) . B(N,K)^T through tensor memory (TS form); device pointers, bf16 operands, fp32 out
long long gru_tc_test_mma(const bf16* A, const bf16* B, float* D, int N, int K, int nacc, cudaStream_t s) {
    if (N % 16 || N < 16 || N > 128 || K % 64 || K < 64 || K > 512) throw std::runtime_error("gru_tc_test_mma: N in 16..128 step 16, K in 64..512 step 64");
    if (nacc < 1 || nacc * N > 256) throw std::runtime_error("gru_tc_test_mma: nacc * N must fit 256 accumulator columns");
    const size_t smem = 1024 + (size_t)N * K * 2 + 64;
    CUDA_CHECK(cudaFuncSetAttribute(k_test_ts_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* dc;
    CUDA_CHECK(cudaMalloc(&dc, 16));
    long long best = 1ll << 60;
    for (int rep = 0; rep < 3; ++rep) {   // the last runs have warm instruction caches
        k_test_ts_mma<<<1, 128, smem, s>>>(A, B, D, N, K, nacc, dc);
        CUDA_CHECK(cudaGetLastError());
        COUNT_LAUNCH();
        long long hc[2] = {0, 0};
        CUDA_CHECK(cudaStreamSynchronize(s));
        CUDA_CHECK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
        if (hc[0] < best) best = hc[0] | (hc[1] << 32);     // low word: issue -> completion, high word: issue loop alone
    }
    cudaFree(dc);
    return best;
}
