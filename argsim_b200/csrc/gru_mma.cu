// gru_mma.cu -- persistent GRU recurrence for H = 512 (bf16 operands, fp32 state and gates).
//
// The recurrence is bound by serial-step latency, not by flops: per step and direction-layer it
// is a (rows x 512) x (512 x 1536) product that cannot start before the previous step finished.
// Design (DESIGN.md "recurrence"):
//   * a GROUP of 16 CTAs owns one (direction, batch-slice); CTA c owns hidden units [32c, 32c+32),
//     i.e. 96 rows of R.  Its R slice (96 x 512 bf16 = 96 KB) is loaded ONCE into registers as
//     mma.sync A-fragments (96 registers per thread) and stays there for the whole sequence --
//     no shared-memory or L2 re-read of weights per step.
//   * forward:  gh[96 x n] = R_own[96 x 512] . h^T[512 x n]  needs the whole h of the previous step:
//     every CTA publishes its 32 new units per row through an L2-resident "LL" buffer (8-byte words =
//     two bf16 + a 32-bit step tag, so data and flag arrive in ONE store: no fence, no separate flag).
//   * backward: dh_prev[n x 512] = dgh[n x 1536] . R is computed by the dual decomposition: CTA c
//     multiplies its OWN 96 dgh columns (local, no gather) with R_own^T[512 x 96] and scatters the
//     512 partial sums to their owner CTAs through the same LL mechanism (reduce-scatter), so the
//     bytes exchanged per step equal the forward's instead of 3x.
//   * rows of a slice are processed in chunks of 16 (two n=8 MMA tiles); fp32 h state lives in smem.
//   * groups are independent (no grid-wide barrier); the launch is cooperative only to guarantee
//     that the 16 CTAs that wait on each other are co-resident.
// Tensor-core instruction: mma.sync.m16n8k16 (bf16 -> fp32).  tcgen05 is the wrong tool here: its
// minimum M = 128 tile and shared-memory operand reads (>= 96 KB per step) cost more than the whole
// register-stationary step; the batched GEMMs around the recurrence use tcgen05 (gemm_tc.cu).
#include "kernels.h"
#include "plan.h"
#include <cooperative_groups.h>
#include <vector>

namespace {

constexpr int HH = 512;
constexpr int CL = 16;
constexpr int UN = 32;
constexpr int NTH = 256;
constexpr int CH = 16;
constexpr int HS_LD = HH + 8;
constexpr int RED_LD = 100;
constexpr int GS_LD = 96 + 8;
constexpr int MAX_BSL = 128;

struct FwdDirP {
    const float* gx; const bf16* R; const float* bR; const float* h0;
    float* hs_f; bf16* hs_h; float* cache;
    float* hT;    // (b,H) fp32 state after the last step of the segment (sorted order) or null
    int ld_gx, ld_hs, reverse;
};
struct FwdP {
    FwdDirP dir[2];
    const int* off; const int* nact;
    unsigned long long* xbuf;
    int ndir, nslices, b, Ttot, t0, Tseg, bslr;   // steps [t0, t0+Tseg) of a plan with Ttot steps
    unsigned tag_base;
    int variant;       // bit0: gx loaded one step ahead into registers; bit1: L2 prefetch two steps ahead by idle warps
    long long* prof;   // debug: per-phase clock totals of thread 0 of every CTA (8 slots each) or null
};
struct BwdDirP {
    const float* dhs; const float* hs_f; const bf16* hs_h; const float* h0; const float* cache; const bf16* R;
    float* dgx_f; bf16* dgx_h; float* dgh_f; bf16* dgh_h; float* hp_f; bf16* hp_h; float* dh0;
    const float* dh_in;   // (b,H) gradient wrt the state AFTER the segment's last step (from the later segment) or null
    float* dh_out;        // (b,H) gradient wrt the state BEFORE the segment's first step, overwritten (t0 > 0) or null
    int ld_dhs, ld_hs, ld_dg, ld_hp, reverse;
};
struct BwdP {
    BwdDirP dir[2];
    const int* off; const int* nact;
    unsigned long long* ybuf;
    int ndir, nslices, b, Ttot, t0, Tseg, bslr;
    unsigned tag_base;
    int variant;       // bit0: cp.async staging one step ahead; bit1: L2 prefetch two steps ahead by idle warps
    long long* prof;
};

__device__ __forceinline__ void ll_store(unsigned long long* p, uint32_t data, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const unsigned long long* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ll_load1(const unsigned long long* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// volatile asm: keeps its program position relative to ldmatrix / other volatile asm
__device__ __forceinline__ float ld_f32(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint32_t bf16_bits(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16(x)); }
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// tanh(x) = 2 sigm(2x) - 1 with the fast exp / divide (abs error ~1e-7; exact limits at +-inf)
__device__ __forceinline__ float tanh_fast(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
__device__ __forceinline__ int slice_rows(int nat, int sl, int ns) { return nat > sl ? (nat - sl + ns - 1) / ns : 0; }
// a peer that never shows up must not hang the GPU: ~2 s, then trap
#define POLL_GUARD(t0) if (clock64() - (t0) > 4000000000LL) __trap()
#define NA(t) s_nact[(t) - P.t0 + 1]
#define OFF(t) s_off[(t) - P.t0 + 1]
#define PROF_MARK(i) do { if (P.prof && tid == 0) { const long long now_ = clock64(); pacc[i] += now_ - plast; plast = now_; } } while (0)

// =========================================================================================
// forward
// =========================================================================================
__global__ void __launch_bounds__(NTH, 1) k_gru_mma_fwd(const __grid_constant__ FwdP P) {
    extern __shared__ __align__(16) unsigned char sm[];
    bf16* Hs = reinterpret_cast<bf16*>(sm);                                   // [bslr][HS_LD]  h_{t-1}, all 512 units
    float* red = reinterpret_cast<float*>(sm + (size_t)P.bslr * HS_LD * 2);  // [4][CH][RED_LD] k-quarter partial sums
    float* hst = red + 4 * CH * RED_LD;                                       // [bslr][UN]     fp32 state of the own units
    int* s_nact = reinterpret_cast<int*>(hst + (size_t)P.bslr * UN);          // [Tseg+2] step tables: a global load per
    int* s_off = s_nact + P.Tseg + 2;                                         // [Tseg+2] step would sit on the critical path
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < P.Tseg + 2; i += NTH) {
        const int tt = P.t0 - 1 + i;
        s_nact[i] = (tt >= 0 && tt < P.Ttot) ? slice_rows(P.nact[tt], blockIdx.x / CL % P.nslices, P.nslices) : 0;
        s_off[i] = (tt >= 0 && tt <= P.Ttot) ? P.off[tt] : 0;
    }
    const int grp = blockIdx.x / CL, c = blockIdx.x % CL;
    const int ns = P.nslices, d = grp / ns, sl = grp % ns;
    const FwdDirP& A = P.dir[d];
    const int kq = warp & 3, mh = warp >> 2, g4 = lane >> 2, q4 = lane & 3;

    // ---- R slice -> registers (A fragments): local row lr = gate*32 + unit  <->  R row gate*H + 32c + unit
    uint32_t a[3][8][4];
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
        const int lr0 = (3 * mh + mt) * 16 + g4, lr1 = lr0 + 8;
        const bf16* r0 = A.R + (size_t)((lr0 >> 5) * HH + UN * c + (lr0 & 31)) * HH;
        const bf16* r1 = A.R + (size_t)((lr1 >> 5) * HH + UN * c + (lr1 & 31)) * HH;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            const int k0 = 128 * kq + 16 * kt + 2 * q4;
            a[mt][kt][0] = *reinterpret_cast<const uint32_t*>(r0 + k0);
            a[mt][kt][1] = *reinterpret_cast<const uint32_t*>(r1 + k0);
            a[mt][kt][2] = *reinterpret_cast<const uint32_t*>(r0 + k0 + 8);
            a[mt][kt][3] = *reinterpret_cast<const uint32_t*>(r1 + k0 + 8);
        }
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
    __syncthreads();

    int na_prev = 0;
    bool first = true;
    bool have_next = false;
    float gxn[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long plast = clock64();
    for (int k = 0; k < P.Tseg; ++k) {
        const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
        const int na = NA(t);
        if (na == 0) {
            if (A.reverse) continue;
            break;
        }
        PROF_MARK(0);
        const long long row_base = OFF(t);
        const bool had_next = have_next;
        // gx of chunk 0 (independent of h) was loaded one step ahead into gxn; otherwise it is loaded at chunk start
        float gx0[2][3];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = warp + 8 * e;
            if (have_next) {
                gx0[e][0] = gxn[e][0]; gx0[e][1] = gxn[e][1]; gx0[e][2] = gxn[e][2];
            } else {
                gx0[e][0] = gx0[e][1] = gx0[e][2] = 0.f;   // loaded at the start of chunk 0 (after the poll)
            }
        }
        // ---------------- h_{prev} of every active row, all 512 units -> Hs
        if (first) {
            for (int i = tid; i < na * (HH / 2); i += NTH) {
                const int jl = i / (HH / 2), kk = (i % (HH / 2)) * 2;
                float v0 = 0.f, v1 = 0.f;
                if (A.h0) {
                    const float* hp = A.h0 + (size_t)(jl * ns + sl) * HH + kk;
                    v0 = hp[0]; v1 = hp[1];
                }
                *reinterpret_cast<__nv_bfloat162*>(Hs + jl * HS_LD + kk) = __floats2bfloat162_rn(v0, v1);
            }
            first = false;
        } else {
            const int npoll = min(na, na_prev);
            const unsigned tag = P.tag_base + (unsigned)(k - 1);
            const unsigned long long* Xr = X + (size_t)((k - 1) & 1) * xpar;
            const int nvec = npoll * (HH / 4);
            constexpr int U = 8;
            for (int base = 0; base < nvec; base += NTH * U) {
                uint4 x[U];
                bool ok;
                const long long t0 = clock64();
                do {
                    ok = true;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int v = base + u * NTH + tid;
                        if (v < nvec) x[u] = ll_load2(Xr + (size_t)(v / (HH / 4)) * (HH / 2) + 2 * (v % (HH / 4)));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int v = base + u * NTH + tid;
                        if (v < nvec && (x[u].y != tag || x[u].w != tag)) ok = false;
                    }
                    POLL_GUARD(t0);
                } while (!ok);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int v = base + u * NTH + tid;
                    if (v < nvec) *reinterpret_cast<uint2*>(Hs + (v / (HH / 4)) * HS_LD + 4 * (v % (HH / 4))) = make_uint2(x[u].x, x[u].z);
                }
            }
            // rows that join at this step (reverse direction): zero initial state
            for (int i = tid; i < (na - npoll) * (HH / 2); i += NTH) {
                const int jl = npoll + i / (HH / 2), kk = (i % (HH / 2)) * 2;
                *reinterpret_cast<uint32_t*>(Hs + jl * HS_LD + kk) = 0u;
            }
        }
        PROF_MARK(1);   // poll + Hs fill (thread 0's share)
        __syncthreads();
        PROF_MARK(2);   // barrier 1 (waiting for the slowest poller)
        const unsigned tagw = P.tag_base + (unsigned)k;
        unsigned long long* Xw = X + (size_t)(k & 1) * xpar;
        for (int ch = 0; ch * CH < na; ++ch) {
            const int nrows = min(CH, na - ch * CH);
            const int ntl = nrows > 8 ? 2 : 1;   // n=8 MMA tiles that hold live rows (warp-uniform)
            // gx of the rows this thread finishes (row n -> warp n & 7; independent of h: issued before the MMAs)
            float gxv[2][3];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = warp + 8 * e;
                gxv[e][0] = gx0[e][0]; gxv[e][1] = gx0[e][1]; gxv[e][2] = gx0[e][2];
                if (!(P.variant & 4) && (ch > 0 || !had_next) && n < nrows) {
                    const float* gp = A.gx + (size_t)(row_base + (long long)(ch * CH + n) * ns + sl) * A.ld_gx + col;
                    gxv[e][0] = gp[0]; gxv[e][1] = gp[HH]; gxv[e][2] = gp[2 * HH];
                }
            }
            float acc[3][2][4];
#pragma unroll
            for (int mt = 0; mt < 3; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
            for (int kt2 = 0; kt2 < 4; ++kt2) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    if (nt >= ntl) continue;
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(b0, b1, b2, b3, Hs + (ch * CH + nt * 8 + (lane & 7)) * HS_LD + 128 * kq + 32 * kt2 + 8 * (lane >> 3));
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) {
                        mma16816(acc[mt][nt], a[mt][2 * kt2], b0, b1);
                        mma16816(acc[mt][nt], a[mt][2 * kt2 + 1], b2, b3);
                    }
                }
            }
            // this step's gx, issued behind the ldmatrix: loads return in order through L1TEX, so an L2-latency load
            // issued in front of them would hold them up; here it completes under the HMMA tail and barrier 2
            if (P.variant & 4) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = warp + 8 * e;
                    if ((ch > 0 || !had_next) && n < nrows) {
                        const float* gp = A.gx + (size_t)(row_base + (long long)(ch * CH + n) * ns + sl) * A.ld_gx + col;
                        gxv[e][0] = ld_f32(gp); gxv[e][1] = ld_f32(gp + HH); gxv[e][2] = ld_f32(gp + 2 * HH);
                    }
                }
            }
            // next step's gx (chunk 0) -> registers.  Loads return IN ORDER through L1TEX, so they are issued only
            // after this step's ldmatrix (they would stall them) and were made L2 hits by the prefetch.global.L2
            // issued two steps ago; they complete under the HMMAs / the partial-sum barrier.
            if (ch == 0) {
                have_next = false;
                if ((P.variant & 1) && k + 1 < P.Tseg) {
                    const int tn = A.reverse ? t - 1 : t + 1;
                    const int nan = NA(tn);
                    if (nan > 0) {
                        have_next = true;
                        const long long rbn = OFF(tn);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int n = warp + 8 * e;
                            gxn[e][0] = gxn[e][1] = gxn[e][2] = 0.f;
                            if (n < nan) {
                                const float* gp = A.gx + (size_t)(rbn + (long long)n * ns + sl) * A.ld_gx + UN * c + lane;
                                gxn[e][0] = ld_f32(gp); gxn[e][1] = ld_f32(gp + HH); gxn[e][2] = ld_f32(gp + 2 * HH);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int mt = 0; mt < 3; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    if (nt >= ntl) continue;
                    float* rp = red + (kq * CH + nt * 8 + 2 * q4) * RED_LD + (3 * mh + mt) * 16 + g4;
                    rp[0] = acc[mt][nt][0]; rp[RED_LD] = acc[mt][nt][1];
                    rp[8] = acc[mt][nt][2]; rp[RED_LD + 8] = acc[mt][nt][3];
                }
            PROF_MARK(3);   // MMA + partial stores
            __syncthreads();
            PROF_MARK(4);   // barrier 2
            // gx rows of the step after next -> L2.  A prefetch costs its issuing warp ~100 cycles, so it is done by
            // the warps that have no row to finish in this chunk (all of them idle otherwise until barrier 3)
            if ((P.variant & 2) && ch == 0 && warp >= nrows && k + 2 < P.Tseg) {
                const int tn = A.reverse ? t - 2 : t + 2;
                const int nan = NA(tn);
                const long long rbn = OFF(tn);
                const int nidle = NTH / 32 - nrows;
                for (int n = warp - nrows; n < nan; n += nidle) {
                    const float* gp = A.gx + (size_t)(rbn + (long long)n * ns + sl) * A.ld_gx + UN * c;
                    if (lane < 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + lane * HH));
                }
            }
            // ---------------- gates: lane = unit, warp -> rows w, w+8
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = warp + 8 * e;
                if (n < nrows) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float* rb = red + (q * CH + n) * RED_LD;
                        s0 += rb[lane]; s1 += rb[32 + lane]; s2 += rb[64 + lane];
                    }
                    const float r = sigm(gxv[e][0] + s0 + bRr);
                    const float z = sigm(gxv[e][1] + s1 + bRu);
                    const float qq = s2 + bRn;
                    const float nn = tanh_fast(gxv[e][2] + r * qq);
                    const int jl = ch * CH + n;
                    const float hp = hst[jl * UN + lane];
                    const float h = (1.f - z) * nn + z * hp;
                    hst[jl * UN + lane] = h;
                    const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                    const __nv_bfloat16 hb16 = __float2bfloat16(h);
                    if (A.hs_h) A.hs_h[row * A.ld_hs + col] = hb16;
                    if (A.hs_f) A.hs_f[row * A.ld_hs + col] = h;
                    if (A.cache) {
                        float* cp = A.cache + row * 4 * HH + col;
                        cp[0] = r; cp[HH] = z; cp[2 * HH] = nn; cp[3 * HH] = qq;
                    }
                    const uint32_t hb = (uint32_t)__bfloat16_as_ushort(hb16);
                    const uint32_t ob = __shfl_down_sync(0xffffffffu, hb, 1);
                    if (!(lane & 1)) ll_store(Xw + (size_t)jl * (HH / 2) + (col >> 1), hb | (ob << 16), tagw);
                }
            }
            PROF_MARK(5);   // gates + stores
            __syncthreads();
            PROF_MARK(6);   // barrier 3
        }
        na_prev = na;
        PROF_MARK(7);
    }
    if (A.hT) {   // state handed to the next time segment of this layer
        __syncthreads();
        for (int i = tid; i < nloc * UN; i += NTH) {
            const int jl = i >> 5, u = i & 31;
            A.hT[(size_t)(jl * ns + sl) * HH + UN * c + u] = hst[i];
        }
    }
    if (P.prof && tid == 0)
        for (int i = 0; i < 8; ++i) P.prof[blockIdx.x * 8 + i] = pacc[i];
}

// =========================================================================================
// forward, v2: the poll IS the operand fetch
// =========================================================================================
// Same decomposition as k_gru_mma_fwd (16 CTAs per group, 96 rows of R register-stationary per CTA), but the serial
// chain of a step is cut from {poll -> smem -> barrier -> ldmatrix -> HMMA -> smem -> barrier -> gates -> barrier} to
// {poll -> HMMA -> smem -> barrier -> gates}:
//   * the LL exchange buffer is laid out in MMA-FRAGMENT order: word index ((row*32 + kt)*4 + q4)*2 + half holds
//     h[row][16kt + 8half + 2q4 .. +1] (bf16x2) + the step tag, so the 16 bytes a lane needs for one k-tile of its
//     B fragment ({b0,tag,b1,tag}) are ONE ld.volatile.v4 straight from L2 into registers -- no shared-memory staging,
//     no ldmatrix, no block barrier between the poll and the MMAs; every warp starts its HMMAs as soon as ITS words
//     arrived;
//   * the 8 warps split K eight ways (64 k each, all 6 m-tiles: 96 registers of R per thread), so no operand word is
//     fetched twice per CTA and the dependent HMMA chain per accumulator is 4 long; the 8 partial sums meet in a
//     double-buffered shared-memory tile, ONE barrier per chunk;
//   * the gate warp publishes the new h (LL store) BEFORE the bookkeeping stores (hs, gate cache).
constexpr int KW = 8;   // k-slices = warps
template <int OPT>
__global__ void __launch_bounds__(NTH, 1) k_gru_mma_fwd2(const __grid_constant__ FwdP P) {
    extern __shared__ __align__(16) unsigned char sm[];
    float* red = reinterpret_cast<float*>(sm);                         // [2][KW][CH][RED_LD] k-slice partial sums
    float* hst = red + 2 * KW * CH * RED_LD;                           // [bslr][UN] fp32 state of the own units
    int* s_nact = reinterpret_cast<int*>(hst + (size_t)P.bslr * UN);   // [Tseg+2] step tables
    int* s_off = s_nact + P.Tseg + 2;
    float* gxs = reinterpret_cast<float*>(s_off + P.Tseg + 2);         // [2][CH][3][UN] next step's gx of chunk 0 (cp.async)
    if ((int)(blockIdx.x / CL) >= P.ndir * P.nslices) return;          // padding CTA: only there for the 128-block placement
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < P.Tseg + 2; i += NTH) {
        const int tt = P.t0 - 1 + i;
        s_nact[i] = (tt >= 0 && tt < P.Ttot) ? slice_rows(P.nact[tt], blockIdx.x / CL % P.nslices, P.nslices) : 0;
        s_off[i] = (tt >= 0 && tt <= P.Ttot) ? P.off[tt] : 0;
    }
    const int grp = blockIdx.x / CL, c = blockIdx.x % CL;
    const int ns = P.nslices, d = grp / ns, sl = grp % ns;
    const FwdDirP& A = P.dir[d];
    const int g4 = lane >> 2, q4 = lane & 3;
    // compile-time A/B switches (OPT): 1 step tables one step ahead, 2 fine profile marks, 4 k-tile-major LL layout,
    // 8 re-poll only the missing words, 16 next step's gx staged in shared memory by cp.async
    constexpr bool rowmaj = !(OPT & 4), repoll_all = !(OPT & 8), stage_gx = (OPT & 16) != 0;
    // OPT & 128: rows of a slice in chunks of 8 instead of 16. A 16-row slice then runs as two one-tile chunks per step
    // whose exchanges alternate: chunk 0's publish travels while chunk 1 computes and vice versa, so no poll waits.
    constexpr int CW = (OPT & 128) ? 8 : CH;

    // ---- R slice -> registers (A fragments): local row lr = gate*32 + unit <-> R row gate*H + 32c + unit; k in [64w, 64w+64)
    uint32_t a[6][4][4];
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) {
        const int lr0 = mt * 16 + g4, lr1 = lr0 + 8;
        const bf16* r0 = A.R + (size_t)((lr0 >> 5) * HH + UN * c + (lr0 & 31)) * HH;
        const bf16* r1 = A.R + (size_t)((lr1 >> 5) * HH + UN * c + (lr1 & 31)) * HH;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const int k0 = 64 * warp + 16 * kt + 2 * q4;
            a[mt][kt][0] = *reinterpret_cast<const uint32_t*>(r0 + k0);
            a[mt][kt][1] = *reinterpret_cast<const uint32_t*>(r1 + k0);
            a[mt][kt][2] = *reinterpret_cast<const uint32_t*>(r0 + k0 + 8);
            a[mt][kt][3] = *reinterpret_cast<const uint32_t*>(r1 + k0 + 8);
        }
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
    __syncthreads();

    int na_prev = 0;
    bool first = true;
    bool staged = false;   // chunk 0 of the current step has its gx in gxs[k & 1] (copied by THIS thread one step ago)
    unsigned rbuf = 0;
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long plast = clock64();
    // step tables are read one step ahead (the shared-memory latency of the lookups is off the serial chain)
    int na_nx = NA(A.reverse ? P.t0 + P.Tseg - 1 : P.t0), off_nx = OFF(A.reverse ? P.t0 + P.Tseg - 1 : P.t0);
    for (int k = 0; k < P.Tseg; ++k) {
        const int t = A.reverse ? P.t0 + P.Tseg - 1 - k : P.t0 + k;
        const int na = (OPT & 1) ? na_nx : NA(t);
        const long long row_base = (OPT & 1) ? off_nx : OFF(t);
        if ((OPT & 1) && k + 1 < P.Tseg) {
            const int tn = A.reverse ? t - 1 : t + 1;
            na_nx = NA(tn); off_nx = OFF(tn);
        }
        if (na == 0) {
            if (A.reverse) continue;
            break;
        }
        PROF_MARK(0);
        const int npoll = first ? 0 : min(na, na_prev);
        const unsigned tag = P.tag_base + (unsigned)(k - 1), tagw = P.tag_base + (unsigned)k;
        const unsigned long long* Xr = X + (size_t)((k - 1) & 1) * xpar;
        unsigned long long* Xw = X + (size_t)(k & 1) * xpar;
        for (int ch = 0; ch * CW < na; ++ch) {
            const int nrows = min(CW, na - ch * CW);
            const int ntl = nrows > 8 ? 2 : 1;   // n=8 MMA tiles that hold live rows (warp-uniform)
            // ---------------- operand fetch: B fragments of h_{t-1} for this warp's 64 k, rows g4 (+8)
            uint32_t b[2][4][2];
            if (first) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int jl = ch * CW + nt * 8 + g4;
                    const float* hp = (A.h0 && nt < ntl && jl < na) ? A.h0 + (size_t)(jl * ns + sl) * HH + 64 * warp + 2 * q4 : nullptr;
#pragma unroll
                    for (int kt = 0; kt < 4; ++kt) {
                        if (hp) {
                            const float2 v0 = *reinterpret_cast<const float2*>(hp + 16 * kt);
                            const float2 v1 = *reinterpret_cast<const float2*>(hp + 16 * kt + 8);
                            b[nt][kt][0] = bf16_bits(v0.x) | (bf16_bits(v0.y) << 16);
                            b[nt][kt][1] = bf16_bits(v1.x) | (bf16_bits(v1.y) << 16);
                        } else {
                            b[nt][kt][0] = b[nt][kt][1] = 0u;
                        }
                    }
                }
            } else {
                bool need[2];
                const unsigned long long* src[2];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int jl = ch * CW + nt * 8 + g4;
                    need[nt] = nt < ntl && jl < npoll;   // rows >= npoll join here with zero state (reverse direction)
                    src[nt] = rowmaj ? Xr + ((size_t)jl * 32 + 4 * warp) * 8 + 2 * q4 : Xr + ((size_t)(4 * warp) * P.bslr + jl) * 8 + 2 * q4;
                }
                uint4 x[2][4];
                bool miss[2][4];   // words still outstanding: only those are re-polled (fewer L1TEX wavefronts per round)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int kt = 0; kt < 4; ++kt) miss[nt][kt] = need[nt];
                bool ok;
                const long long tp0 = clock64();
                const size_t kts = rowmaj ? 8 : (size_t)P.bslr * 8;
                do {
                    ok = true;
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int kt = 0; kt < 4; ++kt)
                            if (miss[nt][kt]) x[nt][kt] = ll_load2(src[nt] + kts * kt);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int kt = 0; kt < 4; ++kt)
                            if (miss[nt][kt]) {
                                const bool m = (x[nt][kt].y != tag || x[nt][kt].w != tag);
                                if (!repoll_all) miss[nt][kt] = m;
                                if (m) ok = false;
                            }
                    POLL_GUARD(tp0);
                } while (!ok);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int kt = 0; kt < 4; ++kt) {
                        b[nt][kt][0] = need[nt] ? x[nt][kt].x : 0u;
                        b[nt][kt][1] = need[nt] ? x[nt][kt].z : 0u;
                    }
            }
            PROF_MARK(1);   // operand fetch (poll)
            // gx of the rows this warp finishes (row n -> warp n & 7): L2 hits (prefetched two steps ago), they
            // complete under the HMMAs and the barrier
            float gxv[2][3];
            const bool from_smem = stage_gx && ch == 0 && staged;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = warp + 8 * e;
                gxv[e][0] = gxv[e][1] = gxv[e][2] = 0.f;
                if (n < nrows && !from_smem) {
                    const float* gp = A.gx + (size_t)(row_base + (long long)(ch * CW + n) * ns + sl) * A.ld_gx + col;
                    gxv[e][0] = ld_f32(gp); gxv[e][1] = ld_f32(gp + HH); gxv[e][2] = ld_f32(gp + 2 * HH);
                }
            }
            float* redw = red + ((size_t)(rbuf * KW + warp) * CH) * RED_LD;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                if (nt >= ntl) continue;
                float acc[6][4];
#pragma unroll
                for (int mt = 0; mt < 6; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;
#pragma unroll
                for (int kt = 0; kt < 4; ++kt)
#pragma unroll
                    for (int mt = 0; mt < 6; ++mt) mma16816(acc[mt], a[mt][kt], b[nt][kt][0], b[nt][kt][1]);
#pragma unroll
                for (int mt = 0; mt < 6; ++mt) {
                    float* rp = redw + (nt * 8 + 2 * q4) * RED_LD + mt * 16 + g4;
                    rp[0] = acc[mt][0]; rp[RED_LD] = acc[mt][1];
                    rp[8] = acc[mt][2]; rp[RED_LD + 8] = acc[mt][3];
                }
            }
            PROF_MARK(2);   // MMA + partial stores
            __syncthreads();
            PROF_MARK(3);   // barrier issue (the wait itself is deferred to the first dependent instruction)
            // ---------------- gates: lane = unit, warp -> rows w, w+8
            const float* redr = red + ((size_t)(rbuf * KW) * CH) * RED_LD;
            if (from_smem) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");   // own copies, issued one step ago
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = warp + 8 * e;
                    if (n < nrows) {
                        const float* gs = gxs + (((k & 1) * CH + n) * 3) * UN + lane;
                        gxv[e][0] = gs[0]; gxv[e][1] = gs[UN]; gxv[e][2] = gs[2 * UN];
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = warp + 8 * e;
                if (n < nrows) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int q = 0; q < KW; ++q) {
                        const float* rb = redr + (q * CH + n) * RED_LD;
                        s0 += rb[lane]; s1 += rb[32 + lane]; s2 += rb[64 + lane];
                    }
                    if ((OPT & 2) && e == 0) PROF_MARK(4);   // barrier wait + partial-sum reads
                    const float r = sigm(gxv[e][0] + s0 + bRr);
                    const float z = sigm(gxv[e][1] + s1 + bRu);
                    const float qq = s2 + bRn;
                    const float nn = tanh_fast(gxv[e][2] + r * qq);
                    const int jl = ch * CW + n;
                    const float hp = hst[jl * UN + lane];
                    const float h = (1.f - z) * nn + z * hp;
                    const __nv_bfloat16 hb16 = __float2bfloat16(h);
                    // publish first: this store is on every peer's critical path
                    const uint32_t hb = (uint32_t)__bfloat16_as_ushort(hb16);
                    const uint32_t ob = __shfl_down_sync(0xffffffffu, hb, 1);
                    if (!(lane & 1))
                        ll_store(Xw + (rowmaj ? ((size_t)jl * 32 + 2 * c + (lane >> 4)) : ((size_t)(2 * c + (lane >> 4)) * P.bslr + jl)) * 8 + ((lane & 7) >> 1) * 2 + ((lane >> 3) & 1), hb | (ob << 16), tagw);
                    if ((OPT & 2) && e == 0) PROF_MARK(5);   // gate math + publish
                    hst[jl * UN + lane] = h;
                    const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                    if (A.hs_h) A.hs_h[row * A.ld_hs + col] = hb16;
                    if (A.hs_f) A.hs_f[row * A.ld_hs + col] = h;
                    if (A.cache) {
                        float* cp = A.cache + row * 4 * HH + col;
                        cp[0] = r; cp[HH] = z; cp[2 * HH] = nn; cp[3 * HH] = qq;
                    }
                }
            }
            PROF_MARK(6);   // bookkeeping stores (+ second row)
            // next step's gx of chunk 0 -> shared memory (cp.async by the thread that will use it: no barrier needed);
            // issued after the publish, it lands while the exchange is in flight (L2 hit: prefetched one step earlier)
            if (stage_gx && ch == 0) {
                staged = false;
                if (k + 1 < P.Tseg) {
                    const int tn = A.reverse ? t - 1 : t + 1;
                    const int nan = NA(tn);
                    if (nan > 0) {
                        staged = true;
                        const long long rbn = OFF(tn);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int n = warp + 8 * e;
                            if (n < nan) {
                                const float* gp = A.gx + (size_t)(rbn + (long long)n * ns + sl) * A.ld_gx + col;
                                float* gd = gxs + ((((k + 1) & 1) * CH + n) * 3) * UN + lane;
                                cp_async4(gd, gp); cp_async4(gd + UN, gp + HH); cp_async4(gd + 2 * UN, gp + 2 * HH);
                            }
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                }
            }
            if ((OPT & 128) && (P.variant & 2) && k + 2 < P.Tseg) {
                // 8-row chunks keep every warp busy: each warp prefetches the gx row it will finish two steps ahead itself
                const int tn = A.reverse ? t - 2 : t + 2;
                const int n = ch * CW + warp;
                if (n < NA(tn) && lane < 3)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.gx + (size_t)(OFF(tn) + (long long)n * ns + sl) * A.ld_gx + UN * c + lane * HH));
            }
            // gx rows of the step after next -> L2, by the warps that have no row to finish in this chunk
            if (!(OPT & 128) && (P.variant & 2) && ch == 0 && warp >= nrows && k + 2 < P.Tseg) {
                const int tn = A.reverse ? t - 2 : t + 2;
                const int nan = NA(tn);
                const long long rbn = OFF(tn);
                const int nidle = NTH / 32 - nrows;
                for (int n = warp - nrows; n < nan; n += nidle) {
                    const float* gp = A.gx + (size_t)(rbn + (long long)n * ns + sl) * A.ld_gx + UN * c;
                    if (lane < 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + lane * HH));
                }
            }
            rbuf ^= 1u;
        }
        na_prev = na;
        first = false;
        PROF_MARK(7);
    }
    if (A.hT) {   // state handed to the next time segment of this layer
        __syncthreads();
        for (int i = tid; i < nloc * UN; i += NTH) {
            const int jl = i >> 5, u = i & 31;
            A.hT[(size_t)(jl * ns + sl) * HH + UN * c + u] = hst[i];
        }
    }
    if (P.prof && tid == 0)
        for (int i = 0; i < 8; ++i) P.prof[blockIdx.x * 8 + i] = pacc[i];
}

// =========================================================================================
// backward (BPTT)
// =========================================================================================
__device__ __forceinline__ size_t yidx(int par, int dest, int src, int pair, int ul, int npair) {
    return ((((size_t)par * CL + dest) * CL + src) * npair + pair) * UN + ul;
}

// CW = rows per chunk: 16 (two n=8 MMA tiles per chunk) or 8 (one tile; a 16-row slice then runs as two chunks per step
// whose reduce-scatter rounds alternate, each one's exchange in flight while the other computes)
template <int CW>
__global__ void __launch_bounds__(NTH, 1) k_gru_mma_bwd(const __grid_constant__ BwdP P) {
    extern __shared__ __align__(16) unsigned char sm[];
    bf16* Gs0 = reinterpret_cast<bf16*>(sm);                              // [2][CH][GS_LD] own dgh columns of the chunk (double
    float* cs = reinterpret_cast<float*>(sm + 2 * CH * GS_LD * 2);       // buffered: no barrier after the MMA) ; [bslr][UN] d*u carried
    int* s_nact = reinterpret_cast<int*>(cs + (size_t)P.bslr * UN);     // [Tmax], [Tmax+1]: step tables in smem
    int* s_off = s_nact + P.Tseg + 2;
    // cp.async landing zone for the NEXT step's gate inputs of chunk 0 (dhs, r, u, n, q: fp32; h_prev: bf16),
    // double buffered by step parity: their HBM latency is paid one step ahead, off the serial chain
    float* stg = reinterpret_cast<float*>(s_off + P.Tseg + 2);   // [2][CH][5][UN]
    bf16* stgh = reinterpret_cast<bf16*>(stg + 2 * CH * 5 * UN);                            // [2][CH][UN]
    if ((int)(blockIdx.x / CL) >= P.ndir * P.nslices) return;          // padding CTA: only there for the 128-block placement
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < P.Tseg + 2; i += NTH) {
        const int tt = P.t0 - 1 + i;
        s_nact[i] = (tt >= 0 && tt < P.Ttot) ? slice_rows(P.nact[tt], blockIdx.x / CL % P.nslices, P.nslices) : 0;
        s_off[i] = (tt >= 0 && tt <= P.Ttot) ? P.off[tt] : 0;
    }
    const int grp = blockIdx.x / CL, c = blockIdx.x % CL;
    const int ns = P.nslices, d = grp / ns, sl = grp % ns;
    const BwdDirP& A = P.dir[d];
    const int g4 = lane >> 2, q4 = lane & 3;
    const int npair = P.bslr / 2;

    // ---- R_own^T -> registers: A[m = out unit o][k = local gate row lr] = R[grow(lr)][o]
    uint32_t a[4][6][4];
#pragma unroll
    for (int kt = 0; kt < 6; ++kt) {
        const int lr0 = 16 * kt + 2 * q4;
        const unsigned short* R0 = reinterpret_cast<const unsigned short*>(A.R) + (size_t)((lr0 >> 5) * HH + UN * c + (lr0 & 31)) * HH;
        const unsigned short* R1 = R0 + HH;                                   // lr0 + 1 (same gate: lr0 is even)
        const int lr8 = lr0 + 8;
        const unsigned short* R8 = reinterpret_cast<const unsigned short*>(A.R) + (size_t)((lr8 >> 5) * HH + UN * c + (lr8 & 31)) * HH;
        const unsigned short* R9 = R8 + HH;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int o0 = 64 * warp + 16 * mt + g4, o1 = o0 + 8;
            a[mt][kt][0] = (uint32_t)R0[o0] | ((uint32_t)R1[o0] << 16);
            a[mt][kt][1] = (uint32_t)R0[o1] | ((uint32_t)R1[o1] << 16);
            a[mt][kt][2] = (uint32_t)R8[o0] | ((uint32_t)R9[o0] << 16);
            a[mt][kt][3] = (uint32_t)R8[o1] | ((uint32_t)R9[o1] << 16);
        }
    }
    const int col = UN * c + lane;
    const int nloc = slice_rows(P.b, sl, ns);
    for (int i = tid; i < nloc * UN; i += NTH)
        cs[i] = A.dh_in ? A.dh_in[(size_t)((i >> 5) * ns + sl) * HH + UN * c + (i & 31)] : 0.f;
    const size_t ypar = (size_t)CL * CL * npair * UN;
    unsigned long long* Y = P.ybuf + (size_t)grp * 2 * ypar;
    __syncthreads();

    int na_prev = 0;
    bool first = true;
    unsigned gbuf = 0;
    bool staged = false;      // chunk 0 of the current step was prefetched into stg[k & 1]
    int k_last = -1;
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
        // h_{prev} source: the step processed BEFORE t in the forward pass
        const int th = A.reverse ? t + 1 : t - 1;
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
        for (int ch = 0; ch * CW < na; ++ch) {
            const int nrows = min(CW, na - ch * CW);
            const int ntl = nrows > 8 ? 2 : 1;
            // ---- reduce-scatter receive: partial sums of R^T.dgh for my unit and rows (w, w+8) from all 16 CTAs.
            // A word carries rows (2p, 2p+1) of one source; row n lives in pair n>>1, half n&1.
            float pin[2] = {0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = warp + 8 * e;
                if (n < nrows && ch * CW + n < ncarry) {
                    const int pair = ch * (CW / 2) + (n >> 1);
                    uint2 w[CL];
                    bool ok;
                    const long long t0 = clock64();
                    const unsigned long long* yb = Y + yidx(parr, c, 0, pair, lane, npair);
                    const size_t ystr = (size_t)npair * UN;   // stride between sources
                    do {
                        ok = true;
#pragma unroll
                        for (int s = 0; s < CL; ++s) w[s] = ll_load1(yb + s * ystr);
#pragma unroll
                        for (int s = 0; s < CL; ++s)
                            if (w[s].y != tagr) ok = false;
                        POLL_GUARD(t0);
                    } while (!ok);
#pragma unroll
                    for (int s = 0; s < CL; ++s) pin[e] += (n & 1) ? bf16_hi(w[s].x) : bf16_lo(w[s].x);
                }
            }
            PROF_MARK(1);   // reduce-scatter receive
            if (ch == 0 && staged) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");   // this step's inputs, issued one step ago by this thread's warp
                __syncwarp();
            }
            bf16* Gs = Gs0 + gbuf * CH * GS_LD;
            float sv[2][5];      // dr, du, dn, dnr, hp of the rows this warp finishes: written to HBM after the send
            size_t srow[2];
            bool slive[2] = {false, false};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = warp + 8 * e;
                if (e >= ntl) continue;
                float dr = 0.f, du = 0.f, dnr = 0.f;
                if (n < nrows) {
                    const int jl = ch * CW + n;
                    const size_t row = (size_t)(row_base + (long long)jl * ns + sl);
                    const float carry = cs[jl * UN + lane] + (jl < ncarry ? pin[e] : 0.f);
                    float dhv, r, z, nn, qq, hp = 0.f;
                    if (ch == 0 && staged) {
                        const float* d = stg + (((k & 1) * CH + n) * 5) * UN + lane;
                        dhv = d[0]; r = d[UN]; z = d[2 * UN]; nn = d[3 * UN]; qq = d[4 * UN];
                        if (jl < nhp) {
                            if (hp_from_h0) hp = A.h0[(size_t)(jl * ns + sl) * HH + col];
                            else hp = __bfloat162float(stgh[((k & 1) * CH + n) * UN + lane]);
                        }
                    } else {
                        dhv = A.dhs[row * A.ld_dhs + col];
                        const float* cp = A.cache + row * 4 * HH + col;
                        r = cp[0]; z = cp[HH]; nn = cp[2 * HH]; qq = cp[3 * HH];
                        if (jl < nhp) {
                            if (hp_from_h0) {
                                hp = A.h0[(size_t)(jl * ns + sl) * HH + col];
                            } else {
                                const size_t rh = (size_t)(hp_base + (long long)jl * ns + sl);
                                hp = A.hs_f ? A.hs_f[rh * A.ld_hs + col] : __bfloat162float(A.hs_h[rh * A.ld_hs + col]);
                            }
                        }
                    }
                    const float dd = carry + dhv;
                    const float dn = dd * (1.f - z) * (1.f - nn * nn);
                    du = dd * (hp - nn) * z * (1.f - z);
                    dr = dn * qq * r * (1.f - r);
                    dnr = dn * r;
                    cs[jl * UN + lane] = dd * z;
                    sv[e][0] = dr; sv[e][1] = du; sv[e][2] = dn; sv[e][3] = dnr; sv[e][4] = hp;
                    srow[e] = row;
                    slive[e] = true;
                }
                bf16* gp = Gs + n * GS_LD + lane;
                gp[0] = __float2bfloat16(dr); gp[32] = __float2bfloat16(du); gp[64] = __float2bfloat16(dnr);
            }
            PROF_MARK(2);   // gate gradients (global loads + stores)
            __syncthreads();
            PROF_MARK(3);   // barrier 1
            // ---- partial^T[512 x 16] = R_own^T[512 x 96] . dgh_own^T[96 x 16]; this warp: out units 64w..64w+63
            float acc[4][2][4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
            for (int kt2 = 0; kt2 < 3; ++kt2) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    if (nt >= ntl) continue;
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(b0, b1, b2, b3, Gs + (nt * 8 + (lane & 7)) * GS_LD + 32 * kt2 + 8 * (lane >> 3));
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        mma16816(acc[mt][nt], a[mt][2 * kt2], b0, b1);
                        mma16816(acc[mt][nt], a[mt][2 * kt2 + 1], b2, b3);
                    }
                }
            }
            PROF_MARK(4);   // MMA
            // ---- reduce-scatter send: (rows 2q, 2q+1) packed as two bf16 + tag, to the owner of each out unit
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    if (nt >= ntl) continue;
                    const int o0 = 64 * warp + 16 * mt + g4, o1 = o0 + 8;
                    const int pair = ch * (CW / 2) + nt * 4 + q4;
                    ll_store(Y + yidx(parw, o0 >> 5, c, pair, o0 & 31, npair), bf16_bits(acc[mt][nt][0]) | (bf16_bits(acc[mt][nt][1]) << 16), tagw);
                    ll_store(Y + yidx(parw, o1 >> 5, c, pair, o1 & 31, npair), bf16_bits(acc[mt][nt][2]) | (bf16_bits(acc[mt][nt][3]) << 16), tagw);
                }
            PROF_MARK(5);   // send
            // ---- bookkeeping, off the serial chain: this step's gate gradients -> HBM (operands of the batched wgrad / dgrad GEMMs)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (!slive[e] || (P.variant & 4)) continue;   // variant bit 2: measurement only (results invalid)
                const float dr = sv[e][0], du = sv[e][1], dn = sv[e][2], dnr = sv[e][3], hp = sv[e][4];
                const size_t row = srow[e];
                const size_t o = row * A.ld_dg + col;
                if (A.dgx_f) { A.dgx_f[o] = dr; A.dgx_f[o + HH] = du; A.dgx_f[o + 2 * HH] = dn; }
                if (A.dgx_h) { A.dgx_h[o] = __float2bfloat16(dr); A.dgx_h[o + HH] = __float2bfloat16(du); A.dgx_h[o + 2 * HH] = __float2bfloat16(dn); }
                if (A.dgh_f) { A.dgh_f[o] = dr; A.dgh_f[o + HH] = du; A.dgh_f[o + 2 * HH] = dnr; }
                if (A.dgh_h) { A.dgh_h[o] = __float2bfloat16(dr); A.dgh_h[o + HH] = __float2bfloat16(du); A.dgh_h[o + 2 * HH] = __float2bfloat16(dnr); }
                if (A.hp_f) A.hp_f[row * A.ld_hp + col] = hp;
                if (A.hp_h) A.hp_h[row * A.ld_hp + col] = __float2bfloat16(hp);
            }
            // ---- prefetch the NEXT step's chunk-0 gate inputs (after the send: off the serial chain)
            bool staged_next = false;
            if ((P.variant & 1) && ch == 0 && k + 1 < P.Tseg && A.hs_h) {
                const int tq = A.reverse ? t + 1 : t - 1;
                const int naq = NA(tq);
                if (naq > 0) {
                    staged_next = true;
                    const int thq = A.reverse ? tq + 1 : tq - 1;
                    const int nhq = (thq >= 0 && thq < P.Ttot) ? min(naq, NA(thq)) : 0;
                    const long long rbq = OFF(tq), rbh = (thq >= 0 && thq <= P.Ttot) ? OFF(thq) : 0;
                    const int pq = (k + 1) & 1;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int n = warp + 8 * e;
                        if (n < naq && n < CW) {
                            const size_t rq = (size_t)(rbq + (long long)n * ns + sl);
                            float* d = stg + ((pq * CH + n) * 5) * UN + lane;
                            cp_async4(d, A.dhs + rq * A.ld_dhs + col);
                            const float* cq = A.cache + rq * 4 * HH + col;
                            cp_async4(d + UN, cq); cp_async4(d + 2 * UN, cq + HH);
                            cp_async4(d + 3 * UN, cq + 2 * HH); cp_async4(d + 4 * UN, cq + 3 * HH);
                            if (n < nhq && !(lane & 1))
                                cp_async4(stgh + (pq * CH + n) * UN + lane, A.hs_h + (size_t)(rbh + (long long)n * ns + sl) * A.ld_hs + col);
                        }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            // gate inputs of the BPTT step after next -> L2 (by the warps without a row in this chunk), so that the
            // cp.async staging issued next step hits L2 and does not hold up the in-order L1TEX queue
            if ((P.variant & 2) && ch == 0 && warp >= nrows && k + 2 < P.Tseg) {
                const int tn = A.reverse ? t + 2 : t - 2;
                const int nan = NA(tn);
                const long long rbn = OFF(tn);
                const int thn = A.reverse ? tn + 1 : tn - 1;
                const long long rbh = (thn >= 0 && thn < P.Ttot) ? OFF(thn) : -1;
                const int nidle = NTH / 32 - nrows;
                for (int n = warp - nrows; n < nan; n += nidle) {
                    const size_t rown = (size_t)(rbn + (long long)n * ns + sl);
                    if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.cache + rown * 4 * HH + UN * c + lane * HH));
                    if (lane == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.dhs + rown * A.ld_dhs + UN * c));
                    if (lane == 5 && rbh >= 0 && A.hs_h && n < NA(thn))
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.hs_h + (size_t)(rbh + (long long)n * ns + sl) * A.ld_hs + UN * c));
                }
            }
            PROF_MARK(6);   // bookkeeping stores + next-step staging
            if (ch == 0) staged = staged_next;
            gbuf ^= 1u;
        }
        na_prev = na;
        first = false;
        k_last = k;
        PROF_MARK(7);
    }
    if (P.prof && tid == 0)
        for (int i = 0; i < 8; ++i) P.prof[blockIdx.x * 8 + i] = pacc[i];
    // ---- gradient wrt the state before the segment's first forward-pass step.  Forward GRU: its BPTT ends at t0 -- the
    // initial state when t0 = 0 (decoder: accumulated into d ex(z)), else handed to the earlier segment.  Reverse GRU
    // (time-segmented encoder): its BPTT ends at t0 + Tseg - 1; the state before that step belongs to the later segment
    // (the reverse GRU starts from zeros at the end of the sequence, so the last segment hands nothing on).
    const bool accum_dh = !A.reverse && P.t0 == 0;
    float* const dh_dst = A.reverse ? ((P.t0 + P.Tseg < P.Ttot) ? A.dh_out : nullptr) : ((P.t0 == 0) ? A.dh0 : A.dh_out);
    if (dh_dst && k_last >= 0) {
        const unsigned tagr = P.tag_base + (unsigned)k_last;
        const int parr = k_last & 1;
        for (int ch = 0; ch * CW < na_prev; ++ch) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = warp + 8 * e;
                const int jl = ch * CW + n;
                if (n >= CW || jl >= na_prev) continue;
                const int pair = ch * (CW / 2) + (n >> 1);
                float pin = 0.f;
#pragma unroll
                for (int s0 = 0; s0 < CL; s0 += 8) {
                    uint2 w[8];
                    bool ok;
                    const long long t0 = clock64();
                    do {
                        ok = true;
#pragma unroll
                        for (int s = 0; s < 8; ++s) w[s] = ll_load1(Y + yidx(parr, c, s0 + s, pair, lane, npair));
#pragma unroll
                        for (int s = 0; s < 8; ++s)
                            if (w[s].y != tagr) ok = false;
                        POLL_GUARD(t0);
                    } while (!ok);
#pragma unroll
                    for (int s = 0; s < 8; ++s) pin += (n & 1) ? bf16_hi(w[s].x) : bf16_lo(w[s].x);
                }
                const size_t di = (size_t)(jl * ns + sl) * HH + col;
                const float val = cs[jl * UN + lane] + pin;
                dh_dst[di] = accum_dh ? dh_dst[di] + val : val;
            }
        }
    }
}

}  // namespace

struct GruMmaCtx {
    int device = 0, num_sms = 148;
    unsigned launch_id = 1;
    static constexpr int NSLOT = 4;
    unsigned long long* xbuf[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    size_t xcap[NSLOT] = {0, 0, 0, 0};
    unsigned long long* ybuf[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    size_t ycap[NSLOT] = {0, 0, 0, 0};
    bool attr_set = false;
    int variant_fwd = 2, variant_bwd = 3;   // bit 1: L2 prefetch two steps ahead by idle warps; bwd bit 0: cp.async staging
    int fwd2_opt = 0;
    int pad_groups = 8;          // launches that have the chip to themselves are padded to 8 groups = 128 blocks: the block
                                 // dispatcher spreads a 128-block grid over all GPCs (2 CTAs of a group per GPC), a small grid
                                 // is packed into one or two GPCs and its exchange round is ~27 % slower (xbench, profiles/)
    bool chunk8 = false;         // ARGSIM_GRU_CHUNK=8: 8-row chunks (one MMA tile each) in both kernels
    bool fwd_v1 = false;         // ARGSIM_GRU_FWD_V1=1: the first forward kernel (smem-staged operands), for A/B runs
    long long* prof = nullptr;   // ARGSIM_GRU_PROF=1: per-phase clocks, printed to stderr after every launch
};

GruMmaCtx* gru_mma_create(int device) {
    GruMmaCtx* c = new GruMmaCtx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_BSL * HS_LD * 2 + 4 * CH * RED_LD * 4 + MAX_BSL * UN * 4 + 2 * 4100 * 4));
    {
        const int fwd2_smem = 2 * KW * CH * RED_LD * 4 + MAX_BSL * UN * 4 + 2 * 4100 * 4 + 2 * CH * 3 * UN * 4;
        CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_fwd2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2_smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_fwd2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2_smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_fwd2<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2_smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_fwd2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2_smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_fwd2<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2_smem));
        if (const char* v = getenv("ARGSIM_GRU_FWD2_OPT")) c->fwd2_opt = atoi(v);
    }
    c->fwd_v1 = getenv("ARGSIM_GRU_FWD_V1") != nullptr;
    if (const char* v = getenv("ARGSIM_GRU_PAD")) c->pad_groups = atoi(v);
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_bwd<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CH * GS_LD * 2 + MAX_BSL * UN * 4 + 2 * 4100 * 4 + 2 * CH * 5 * UN * 4 + 2 * CH * UN * 2));
    CUDA_CHECK(cudaFuncSetAttribute(k_gru_mma_bwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CH * GS_LD * 2 + MAX_BSL * UN * 4 + 2 * 4100 * 4 + 2 * CH * 5 * UN * 4 + 2 * CH * UN * 2));
    // ARGSIM_GRU_CHUNK=8 (A/B, not measured yet): both kernels walk a slice in chunks of 8 rows (see k_gru_mma_bwd)
    if (const char* v = getenv("ARGSIM_GRU_CHUNK")) c->chunk8 = atoi(v) == 8;
    if (c->chunk8 && c->fwd2_opt == 0) c->fwd2_opt = 128;
    if (const char* v = getenv("ARGSIM_GRU_VARIANT")) { c->variant_fwd = atoi(v) & 7; c->variant_bwd = (atoi(v) >> 3) & 7; }
    if (const char* v = getenv("ARGSIM_GRU_FWD2_VARIANT")) c->variant_fwd = atoi(v);
    if (getenv("ARGSIM_GRU_PROF")) CUDA_CHECK(cudaMalloc(&c->prof, 160 * 8 * sizeof(long long)));
    return c;
}
void gru_mma_destroy(GruMmaCtx* c) {
    if (!c) return;
    for (int i = 0; i < GruMmaCtx::NSLOT; ++i) {
        cudaFree(c->xbuf[i]);
        cudaFree(c->ybuf[i]);
    }
    cudaFree(c->prof);
    delete c;
}
bool gru_mma_supported(int H) { return H == HH; }

static void dump_prof(GruMmaCtx* c, const char* what, int nblocks, int steps, cudaStream_t s) {
    if (!c->prof) return;
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<long long> h(nblocks * 8);
    CUDA_CHECK(cudaMemcpy(h.data(), c->prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg[8] = {0};
    for (int b = 0; b < nblocks; ++b)
        for (int i = 0; i < 8; ++i) avg[i] += (double)h[b * 8 + i] / nblocks;
    fprintf(stderr, "[gru_prof] %s blocks=%d steps=%d cycles/step:", what, nblocks, steps);
    for (int i = 0; i < 8; ++i) fprintf(stderr, " p%d=%.0f", i, avg[i] / steps);
    fprintf(stderr, "  | block0:");
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %.0f", (double)h[i] / steps);
    fprintf(stderr, "\n");
}

// rows_per_slice = 16 fills both n=8 MMA tiles of a chunk; 8 (a launch that has the chip to itself) keeps every slice at
// one MMA tile per step -- about 2,700 instead of 4,700 cycles per step while more than 8 rows per slice are alive
static int small8_rows() {   // single-direction launches with at most this many live rows use 8-row slices (A/B: ARGSIM_SMALL8_ROWS)
    static const int v = getenv("ARGSIM_SMALL8_ROWS") ? atoi(getenv("ARGSIM_SMALL8_ROWS")) : CH;
    return v;
}
static void pick_slices(const GruMmaCtx* c, int ndir, int b, int* ns, int* bslr, int rows_per_slice = CH) {
    const int max_groups = std::max(1, c->num_sms / CL);
    int s = std::max(1, std::min(max_groups / ndir, (b + rows_per_slice - 1) / rows_per_slice));
    int per = (b + s - 1) / s;
    *ns = s;
    *bslr = (per + CH - 1) / CH * CH;
}
bool gru_mma_fits(const GruMmaCtx* c, int ndir, int b) {
    int ns, bslr;
    pick_slices(c, ndir, b, &ns, &bslr);
    return ndir * CL <= c->num_sms && bslr <= MAX_BSL;
}

// One launch = steps [t0, t0+Tseg) of the plan (Tseg < 0: all).  `slot` selects the LL exchange buffer: launches
// that may run CONCURRENTLY (decoder wavefront: one stream per layer) must use different slots.
void gru_mma_fwd(GruMmaCtx* c, const GruFwdArgs* dirs, int ndir, const SeqPlan& Pl, const int* d_off, const int* d_nact, int H,
                 cudaStream_t s, int t0, int Tseg, int slot, int alone, int pad) {
    if (H != HH) throw std::runtime_error("gru_mma: H must be 512");
    if (Tseg < 0) { t0 = 0; Tseg = Pl.Tmax; }
    if (Tseg >= 4096) throw std::runtime_error("gru_mma: more than 4095 steps per launch");
    if (slot < 0 || slot >= GruMmaCtx::NSLOT) throw std::runtime_error("gru_mma: bad slot");
    FwdP P;
    int ns, bslr;
    // sequences alive in the segment: forward directions shrink with t, so the first step has the most
    // rows alive anywhere in [t0, t0+Tseg): nact is non-increasing in t, so the segment's first time step has the most
    // (for either direction of a single-direction launch; a two-direction launch spans all rows)
    const int b_seg = (ndir == 1) ? Pl.nact[t0] : Pl.b;
    pick_slices(c, ndir, b_seg, &ns, &bslr, (alone == 1 || (alone == 0 && ndir == 1 && b_seg <= small8_rows())) ? 8 : CH);   // alone: 0 = by live rows (<= 16: two groups of <= 8 rows cost 16 more SMs and halve the MMA work per step), 1 = 8-row slices, 2 = 16-row slices
    if (bslr > MAX_BSL) throw std::runtime_error("gru_mma: batch too large for the persistent kernel");
    for (int d = 0; d < ndir; ++d) {
        const GruFwdArgs& a = dirs[d];
        if (!a.R_h) throw std::runtime_error("gru_mma: bf16 weights missing");
        if (a.h0 && a.reverse && ndir != 1) throw std::runtime_error("gru_mma: a reverse direction takes h0 only in a single-direction (segment) launch");
        P.dir[d] = FwdDirP{a.gx, a.R_h, a.bR, a.h0, a.hs_f, a.hs_h, a.cache, a.hT, a.ld_gx, a.ld_hs, a.reverse};
    }
    if (ndir == 1) P.dir[1] = P.dir[0];
    const int groups = ndir * ns;
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
    P.prof = c->prof;
    P.variant = c->variant_fwd;
    if (c->launch_id >= (1u << 20)) c->launch_id = 1;
    const size_t smem = c->fwd_v1 ? (size_t)bslr * HS_LD * 2 + 4 * CH * RED_LD * 4 + (size_t)bslr * UN * 4 + (size_t)(2 * Tseg + 4) * 4
                                  : (size_t)2 * KW * CH * RED_LD * 4 + (size_t)bslr * UN * 4 + (size_t)(2 * Tseg + 4) * 4 + 2 * CH * 3 * UN * 4;
    void* args[] = {&P};
    void* fwd2_fn = c->fwd2_opt == 2 ? (void*)k_gru_mma_fwd2<2> : c->fwd2_opt == 12 ? (void*)k_gru_mma_fwd2<12>
                  : c->fwd2_opt == 16 ? (void*)k_gru_mma_fwd2<16> : c->fwd2_opt == 128 ? (void*)k_gru_mma_fwd2<128>
                  : (void*)k_gru_mma_fwd2<0>;
    const int grid_groups = (!c->fwd_v1 && pad) ? std::max(groups, c->pad_groups) : groups;
    if (pad == 2 && !c->fwd_v1) {
        // padded launch next to other recurrence launches (wavefront / segment chains): a cooperative launch would wait for
        // 128 free SMs.  A plain launch is safe because the padding blocks exit at once and the caller's slice budget keeps
        // the ACTIVE blocks of all launches that can be in flight together within the SM count (1 CTA per SM): every
        // active block gets an SM without anybody having to finish first.
        CUDA_CHECK(cudaLaunchKernel(fwd2_fn, dim3(grid_groups * CL), dim3(NTH), args, smem, s));
    } else
    CUDA_CHECK(cudaLaunchCooperativeKernel(c->fwd_v1 ? (void*)k_gru_mma_fwd : fwd2_fn, dim3(grid_groups * CL), dim3(NTH), args, smem, s));
    COUNT_LAUNCH();
    dump_prof(c, ndir == 2 ? "fwd_enc" : "fwd_dec", groups * CL, Tseg, s);
}

void gru_mma_bwd(GruMmaCtx* c, const GruBwdArgs* dirs, int ndir, const SeqPlan& Pl, const int* d_off, const int* d_nact, int H,
                 cudaStream_t s, int t0, int Tseg, int slot, int alone, int pad, int chunk) {
    if (H != HH) throw std::runtime_error("gru_mma: H must be 512");
    if (Tseg < 0) { t0 = 0; Tseg = Pl.Tmax; }
    if (Tseg >= 4096) throw std::runtime_error("gru_mma: more than 4095 steps per launch");
    if (slot < 0 || slot >= GruMmaCtx::NSLOT) throw std::runtime_error("gru_mma: bad slot");
    BwdP P;
    int ns, bslr;
    const int b_seg = (ndir == 1) ? Pl.nact[t0] : Pl.b;
    pick_slices(c, ndir, b_seg, &ns, &bslr, (alone == 1 || (alone == 0 && ndir == 1 && b_seg <= small8_rows())) ? 8 : CH);   // alone: 0 = by live rows (<= 16: two groups of <= 8 rows cost 16 more SMs and halve the MMA work per step), 1 = 8-row slices, 2 = 16-row slices
    if (bslr > MAX_BSL) throw std::runtime_error("gru_mma: batch too large for the persistent kernel");
    for (int d = 0; d < ndir; ++d) {
        const GruBwdArgs& a = dirs[d];
        if (!a.R_h) throw std::runtime_error("gru_mma: bf16 weights missing");
        P.dir[d] = BwdDirP{a.dhs, a.hs_f, a.hs_h, a.h0, a.cache, a.R_h, a.dgx_f, a.dgx_h, a.dgh_f, a.dgh_h, a.hp_f, a.hp_h, a.dh0,
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
    P.prof = c->prof;
    P.variant = c->variant_bwd;
    if (c->launch_id >= (1u << 20)) c->launch_id = 1;
    const size_t smem = (size_t)2 * CH * GS_LD * 2 + (size_t)bslr * UN * 4 + (size_t)(2 * Tseg + 4) * 4 + 2 * CH * 5 * UN * 4 + 2 * CH * UN * 2;
    void* args[] = {&P};
    const int grid_groups = pad ? std::max(groups, c->pad_groups) : groups;
    void* bwd_fn = (chunk == 8 || (chunk == 0 && c->chunk8)) ? (void*)k_gru_mma_bwd<8> : (void*)k_gru_mma_bwd<CH>;
    if (pad == 2) CUDA_CHECK(cudaLaunchKernel(bwd_fn, dim3(grid_groups * CL), dim3(NTH), args, smem, s));   // see gru_mma_fwd
    else
    CUDA_CHECK(cudaLaunchCooperativeKernel(bwd_fn, dim3(grid_groups * CL), dim3(NTH), args, smem, s));
    COUNT_LAUNCH();
    dump_prof(c, ndir == 2 ? "bwd_enc" : "bwd_dec", groups * CL, Tseg, s);
}
