// gru_generic.cu -- reference-order GRU recurrence, one GEMM + one gate kernel per time step.
// The per-step GEMM is fp32 SIMT in the validation mode and, in bf16 mode with a hidden size the persistent
// kernel does not cover (H != 512, e.g. the 4x-wide BASELINE configs[4]), the tcgen05 GEMM on the bf16 weight
// shadow (split-K through TMA reduce-add: R no longer fits on chip, it streams from L2 every step).
// Used by the FP32_VALIDATE precision mode (parity runs against the fp64/fp32 oracle at 1e-3),
// for hidden sizes the persistent kernel does not cover, and as the on-device checker of
// gru_mma.cu.  Cell = cuDNN form (src/model.py:15):
//   r = sig(gx_r + R_r h + bR_r)   u = sig(gx_u + R_u h + bR_u)
//   n = tanh(gx_n + r * (R_n h + bR_n))   h' = (1-u) n + u h
#include "kernels.h"
#include "plan.h"

namespace {

// nzero > na: the launch also covers rows [na, nzero) and clears their (dead) gh entries; every gh entry read is cleared
// too, so the next step's split-K GEMM can reduce-add into gh without a memset in between (one device op less per step)
__global__ void __launch_bounds__(256) k_gate_fwd(const float* __restrict__ gx, int ld_gx, float* __restrict__ gh,
                                                  const float* __restrict__ bR, float* __restrict__ state,
                                                  float* __restrict__ hs_f, bf16* __restrict__ hs_h, int ld_hs,
                                                  float* __restrict__ cache, int na, int H, bf16* __restrict__ state_h, int nzero) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= max(na, nzero) * H) return;
    const int j = idx / H, u = idx - j * H;
    float* q3 = gh + (long long)j * 3 * H;
    if (j >= na) {
        q3[u] = 0.f; q3[H + u] = 0.f; q3[2 * H + u] = 0.f;
        return;
    }
    const float* g = gx + (long long)j * ld_gx;
    const float ghr = q3[u], ghu = q3[H + u], ghn = q3[2 * H + u];
    if (nzero) { q3[u] = 0.f; q3[H + u] = 0.f; q3[2 * H + u] = 0.f; }
    const float r = 1.f / (1.f + expf(-(g[u] + ghr + bR[u])));
    const float z = 1.f / (1.f + expf(-(g[H + u] + ghu + bR[H + u])));
    const float q = ghn + bR[2 * H + u];
    const float n = tanhf(g[2 * H + u] + r * q);
    const float hp = state[(long long)j * H + u];
    const float h = (1.f - z) * n + z * hp;
    state[(long long)j * H + u] = h;
    if (state_h) state_h[(long long)j * H + u] = __float2bfloat16(h);
    if (hs_f) hs_f[(long long)j * ld_hs + u] = h;
    if (hs_h) hs_h[(long long)j * ld_hs + u] = __float2bfloat16(h);
    if (cache) {
        float* c = cache + (long long)j * 4 * H;
        c[u] = r; c[H + u] = z; c[2 * H + u] = n; c[3 * H + u] = q;
    }
}

// hp_src_* point at the row of sequence 0 of the previous step (or null -> h0 / zeros); nprev rows valid
__global__ void __launch_bounds__(256) k_gate_bwd(const float* __restrict__ dhs, int ld_dhs, const float* __restrict__ cache,
                                                  const float* __restrict__ hp_f, const bf16* __restrict__ hp_h, int ld_hp_src,
                                                  int nprev, const float* __restrict__ h0, float* __restrict__ carry,
                                                  float* __restrict__ tmp_dgh, float* __restrict__ dgx_f,
                                                  bf16* __restrict__ dgx_h, float* __restrict__ dgh_f, bf16* __restrict__ dgh_h,
                                                  int ld_dg, float* __restrict__ hpo_f, bf16* __restrict__ hpo_h, int ld_hpo,
                                                  int na, int H) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    // programmatic dependent launch (no-ops on a plain launch): the GEMM behind this kernel may become resident; the carry this
    // kernel reads is complete after the wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (idx >= na * H) return;
    const int j = idx / H, u = idx - j * H;
    const float* c = cache + (long long)j * 4 * H;
    const float r = c[u], z = c[H + u], n = c[2 * H + u], q = c[3 * H + u];
    float hp = 0.f;
    if (j < nprev) {
        if (hp_f) hp = hp_f[(long long)j * ld_hp_src + u];
        else if (hp_h) hp = __bfloat162float(hp_h[(long long)j * ld_hp_src + u]);
        else if (h0) hp = h0[(long long)j * H + u];
    }
    const float d = carry[(long long)j * H + u] + dhs[(long long)j * ld_dhs + u];
    const float dn = d * (1.f - z) * (1.f - n * n);
    const float du = d * (hp - n) * z * (1.f - z);
    const float dr = dn * q * r * (1.f - r);
    const float dnr = dn * r;
    carry[(long long)j * H + u] = d * z;
    float* t3 = tmp_dgh + (long long)j * 3 * H;
    t3[u] = dr; t3[H + u] = du; t3[2 * H + u] = dnr;
    const long long o = (long long)j * ld_dg;
    if (dgx_f) { dgx_f[o + u] = dr; dgx_f[o + H + u] = du; dgx_f[o + 2 * H + u] = dn; }
    if (dgx_h) { dgx_h[o + u] = __float2bfloat16(dr); dgx_h[o + H + u] = __float2bfloat16(du); dgx_h[o + 2 * H + u] = __float2bfloat16(dn); }
    if (dgh_f) { dgh_f[o + u] = dr; dgh_f[o + H + u] = du; dgh_f[o + 2 * H + u] = dnr; }
    if (dgh_h) { dgh_h[o + u] = __float2bfloat16(dr); dgh_h[o + H + u] = __float2bfloat16(du); dgh_h[o + 2 * H + u] = __float2bfloat16(dnr); }
    if (hpo_f) hpo_f[(long long)j * ld_hpo + u] = hp;
    if (hpo_h) hpo_h[(long long)j * ld_hpo + u] = __float2bfloat16(hp);
}

__global__ void __launch_bounds__(256) k_axpy(float* __restrict__ y, const float* __restrict__ x, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

cudaEvent_t g_ev[4];
bool g_ev_init = false;
void fork_join_init() {
    if (g_ev_init) return;
    for (auto& e : g_ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_ev_init = true;
}
}  // namespace

__global__ void __launch_bounds__(256) k_state_to_bf16(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2bfloat16(x[i]);
}
static bool use_tc_gemm(const bf16* R_h, int H) {
    static const bool off = getenv("ARGSIM_GENERIC_SIMT") != nullptr;
    return !off && R_h != nullptr && gemm_tc_available() && H % 64 == 0;
}

// per direction: fp32 state / carry (b,H), gh / dgh scratch (b,3H), bf16 state (b,H)
size_t gru_generic_work_floats(int b, int H) { return (size_t)b * H * 5; }

void gru_generic_fwd(const GruFwdArgs* dirs, int ndir, const SeqPlan& P, int H, float* work, cudaStream_t* streams) {
    fork_join_init();
    const size_t wstride = gru_generic_work_floats(P.b, H);
    // bf16 mode: ONE fused kernel per time step for all directions (TMA ring + tcgen05 + gate epilogue, gru_tc.cu) instead of
    // {memset, split-K GEMM, gate kernel} per direction (ARGSIM_GENERIC_UNFUSED=1 restores the latter)
    static const bool unfused = getenv("ARGSIM_GENERIC_UNFUSED") != nullptr;
    if (!unfused && ndir <= 2 && use_tc_gemm(dirs[0].R_h, H) && gru_step_supported(H, P.b)) {
        cudaStream_t s = streams[0];
        float* state[2];
        bf16* sh[2][2];
        for (int d = 0; d < ndir; ++d) {
            const GruFwdArgs& a = dirs[d];
            state[d] = work + d * wstride;
            sh[d][0] = reinterpret_cast<bf16*>(state[d] + (size_t)P.b * 4 * H);      // two bf16 (b,H) buffers in the last b*H floats
            sh[d][1] = sh[d][0] + (size_t)P.b * H;
            if (a.h0 && !a.reverse) CUDA_CHECK(cudaMemcpyAsync(state[d], a.h0, sizeof(float) * P.b * H, cudaMemcpyDeviceToDevice, s));
            else CUDA_CHECK(cudaMemsetAsync(state[d], 0, sizeof(float) * P.b * H, s));
            k_state_to_bf16<<<cdiv((long long)P.b * H, 256), 256, 0, s>>>(state[d], sh[d][0], (long long)P.b * H);
            COUNT_LAUNCH();
            CUDA_CHECK(cudaMemcpyAsync(sh[d][1], sh[d][0], sizeof(bf16) * P.b * H, cudaMemcpyDeviceToDevice, s));
        }
        for (int k = 0; k < P.Tmax; ++k) {
            int na[2] = {0, 0};
            long long row0[2] = {0, 0};
            bf16 *cur[2], *nxt[2];
            for (int d = 0; d < ndir; ++d) {
                const int t = dirs[d].reverse ? P.Tmax - 1 - k : k;
                na[d] = P.nact[t];
                row0[d] = P.off[t];
                cur[d] = sh[d][k & 1];
                nxt[d] = sh[d][(k + 1) & 1];
            }
            gru_step_fwd(dirs, ndir, na, row0, P.b, H, state, cur, nxt, s);
        }
        return;
    }
    if (ndir > 1) {
        CUDA_CHECK(cudaEventRecord(g_ev[0], streams[0]));
        for (int d = 1; d < ndir; ++d) CUDA_CHECK(cudaStreamWaitEvent(streams[d], g_ev[0], 0));
    }
    for (int d = 0; d < ndir; ++d) {
        const GruFwdArgs& a = dirs[d];
        cudaStream_t s = streams[d];
        float* state = work + d * wstride;
        float* gh = state + (size_t)P.b * H;
        bf16* state_h = use_tc_gemm(a.R_h, H) ? reinterpret_cast<bf16*>(gh + (size_t)P.b * 3 * H) : nullptr;
        if (a.h0 && !a.reverse)
            CUDA_CHECK(cudaMemcpyAsync(state, a.h0, sizeof(float) * P.b * H, cudaMemcpyDeviceToDevice, s));
        else
            CUDA_CHECK(cudaMemsetAsync(state, 0, sizeof(float) * P.b * H, s));
        // tcgen05 path with a batch that fits one 128-row tile: the GEMM covers all P.b rows with constant shapes (cached tensor
        // maps) and reduce-adds into a gh that the gate kernel clears behind itself
        const bool selfclear = state_h != nullptr && P.b <= 128;
        if (state_h) {
            k_state_to_bf16<<<cdiv((long long)P.b * H, 256), 256, 0, s>>>(state, state_h, (long long)P.b * H);
            COUNT_LAUNCH();
        }
        if (selfclear) CUDA_CHECK(cudaMemsetAsync(gh, 0, sizeof(float) * P.b * 3 * H, s));
        for (int k = 0; k < P.Tmax; ++k) {
            const int t = a.reverse ? P.Tmax - 1 - k : k;
            const int na = P.nact[t];
            const long long r0 = P.off[t];
            // all P.b state rows every step (rows >= na are dead and ignored by the gate kernel): an M = 64 product fills the same
            // 128-row MMA tile as M = na, and constant shapes keep the three tensor maps in gemm_tc's cache
            if (selfclear) gemm_tc(state_h, H, 0, a.R_h, H, 0, gh, nullptr, 3 * H, P.b, 3 * H, H, 1.f, nullptr, /*accumulate*/ 1, s);
            else if (state_h) gemm_tc(state_h, H, 0, a.R_h, H, 0, gh, nullptr, 3 * H, na, 3 * H, H, 1.f, nullptr, 0, s);
            else gemm_simt(state, H, 0, a.R_f, H, 0, gh, 3 * H, na, 3 * H, H, 1.f, nullptr, 0, nullptr, s);
            k_gate_fwd<<<cdiv((long long)(selfclear ? P.b : na) * H, 256), 256, 0, s>>>(
                a.gx + r0 * a.ld_gx, a.ld_gx, gh, a.bR, state, a.hs_f ? a.hs_f + r0 * a.ld_hs : nullptr,
                a.hs_h ? a.hs_h + r0 * a.ld_hs : nullptr, a.ld_hs, a.cache ? a.cache + r0 * 4 * H : nullptr, na, H, state_h,
                selfclear ? P.b : 0);
            COUNT_LAUNCH();
        }
    }
    if (ndir > 1) {
        for (int d = 1; d < ndir; ++d) {
            CUDA_CHECK(cudaEventRecord(g_ev[d], streams[d]));
            CUDA_CHECK(cudaStreamWaitEvent(streams[0], g_ev[d], 0));
        }
    }
}

void gru_generic_bwd(const GruBwdArgs* dirs, int ndir, const SeqPlan& P, int H, float* work, cudaStream_t* streams) {
    // per-step chain {gate gradients, carry += dgh.R}: programmatic dependent launch hides each launch under its predecessor
    static const bool pdl = getenv("ARGSIM_GENERIC_NO_PDL") == nullptr;
    fork_join_init();
    const size_t wstride = gru_generic_work_floats(P.b, H);
    if (ndir > 1) {
        CUDA_CHECK(cudaEventRecord(g_ev[0], streams[0]));
        for (int d = 1; d < ndir; ++d) CUDA_CHECK(cudaStreamWaitEvent(streams[d], g_ev[0], 0));
    }
    for (int d = 0; d < ndir; ++d) {
        const GruBwdArgs& a = dirs[d];
        cudaStream_t s = streams[d];
        float* carry = work + d * wstride;
        float* tmp = carry + (size_t)P.b * H;
        CUDA_CHECK(cudaMemsetAsync(carry, 0, sizeof(float) * P.b * H, s));
        for (int k = P.Tmax - 1; k >= 0; --k) {  // reverse of the forward processing order
            const int t = a.reverse ? P.Tmax - 1 - k : k;
            const int na = P.nact[t];
            const long long r0 = P.off[t];
            // previous (in forward processing order) step's outputs = h_prev of this step
            const float* hp_f = nullptr;
            const bf16* hp_h = nullptr;
            const float* h0 = nullptr;
            int nprev = 0;
            const int tp = a.reverse ? t + 1 : t - 1;
            if (tp >= 0 && tp < P.Tmax) {
                nprev = std::min(na, P.nact[tp]);
                if (a.hs_f) hp_f = a.hs_f + (long long)P.off[tp] * a.ld_hs;
                else hp_h = a.hs_h + (long long)P.off[tp] * a.ld_hs;
            } else if (!a.reverse && a.h0) {
                h0 = a.h0;
                nprev = na;
            }
            const bool tc_step = a.dgh_h && a.ld_dg % 8 == 0 && use_tc_gemm(a.R_h, H);
            {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)cdiv((long long)na * H, 256)); cfg.blockDim = dim3(256); cfg.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = (tc_step && pdl) ? 1 : 0;
                CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gate_bwd,
                    a.dhs + r0 * a.ld_dhs, a.ld_dhs, a.cache + r0 * 4 * H, hp_f, hp_h, a.ld_hs, nprev, h0, carry, tmp,
                    a.dgx_f ? a.dgx_f + r0 * a.ld_dg : nullptr, a.dgx_h ? a.dgx_h + r0 * a.ld_dg : nullptr,
                    a.dgh_f ? a.dgh_f + r0 * a.ld_dg : nullptr, a.dgh_h ? a.dgh_h + r0 * a.ld_dg : nullptr, a.ld_dg,
                    a.hp_f ? a.hp_f + r0 * a.ld_hp : nullptr, a.hp_h ? a.hp_h + r0 * a.ld_hp : nullptr, a.ld_hp, na, H));
            }
            COUNT_LAUNCH();
            // carry[0:na] += dgh[0:na] . R        (R is (3H,H): stored (K, N) -> b_mn = 1)
            if (tc_step) {
                gemm_tc_pdl(pdl);
                gemm_tc(a.dgh_h + r0 * a.ld_dg, a.ld_dg, 0, a.R_h, H, 1, carry, nullptr, H, na, H, 3 * H, 1.f, nullptr, 1, s);
                gemm_tc_pdl(false);
            } else
                gemm_simt(tmp, 3 * H, 0, a.R_f, H, 1, carry, H, na, H, 3 * H, 1.f, nullptr, 1, nullptr, s);
        }
        if (a.dh0 && !a.reverse) {
            k_axpy<<<cdiv((long long)P.b * H, 256), 256, 0, s>>>(a.dh0, carry, (long long)P.b * H);
            COUNT_LAUNCH();
        }
    }
    if (ndir > 1) {
        for (int d = 1; d < ndir; ++d) {
            CUDA_CHECK(cudaEventRecord(g_ev[d], streams[d]));
            CUDA_CHECK(cudaStreamWaitEvent(streams[0], g_ev[d], 0));
        }
    }
}

void gru_generic_cell(const float* gx, int ld_gx, const float* R, const float* bR, float* state, float* gh_work, int nb,
                      int H, cudaStream_t s) {
    gemm_simt(state, H, 0, R, H, 0, gh_work, 3 * H, nb, 3 * H, H, 1.f, nullptr, 0, nullptr, s);
    k_gate_fwd<<<cdiv((long long)nb * H, 256), 256, 0, s>>>(gx, ld_gx, gh_work, bR, state, nullptr, nullptr, H, nullptr, nb, H, nullptr, 0);
    COUNT_LAUNCH();
}
