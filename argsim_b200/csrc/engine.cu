// engine.cu -- see engine.h.  Reference: src/model.py:75-189 (graph), src/train.py:104-121 (loops).
#include "engine.h"
#include <chrono>
#include "nccl_dyn.h"
#include <nvtx3/nvToolsExt.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <random>

#define RUN(...)                  \
    do {                          \
        if (!arena.dry) {         \
            __VA_ARGS__;          \
        }                         \
    } while (0)

static inline size_t align_up(size_t n, size_t a) { return (n + a - 1) / a * a; }

// Blocking copy ordered on the engine's own stream.  st[0] is created cudaStreamNonBlocking, so the
// legacy default stream used by plain cudaMemcpy is NOT ordered against the kernels launched on it
// (a pageable H2D cudaMemcpy may even return before its DMA lands).
void Engine::copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, kind, st[0]));
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
}

// ------------------------------------------------------------------------------------------
// construction: parameter table in backward-completion order (== allreduce bucket order)
// ------------------------------------------------------------------------------------------
Engine::Engine(const argsim_config& c) : cfg(c) {
    V = c.dim_tgt; D = c.dim_emb; R = c.dim_rep; L = c.rnn_layers; H = D;
    if (V <= 0 || D <= 0 || R <= 0 || L <= 0) throw std::runtime_error("bad model dimensions");
    attentive = c.attentive != 0;   // src/model.py:136-145 ('todo fixme' there; the repaired form is stated in DESIGN.md section 6)
    enc_kind = (c.bidirectional && c.bidir_stacked) ? 0 : (c.bidirectional ? 1 : 2);
    EH = (enc_kind == 2) ? H : 2 * H;
    tied = c.logit_use_embed != 0;   // false: separate (D,V) projection + bias (src/model.py:167-168)
    if (D % 8 || R % 8 || V % 8) throw std::runtime_error("dim_tgt, dim_emb and dim_rep must be multiples of 8");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw std::runtime_error("no CUDA device: argsim_b200 has no CPU fallback");
    CUDA_CHECK(cudaSetDevice(c.device));
    is_bf16 = (c.precision == ARGSIM_BF16);
    use_tc = is_bf16;
    if (use_tc) {
        gemm_tc_init(c.device);
        if (!gemm_tc_available()) throw std::runtime_error("BF16 precision needs the tcgen05 GEMM path (sm_100a device + driver TMA entry point)");
    }
    use_mma = is_bf16 && gru_mma_supported(H) && !(c.flags & 4);
    // the serial chain (main stream, recurrence chains) outranks the side stream that carries the weight-gradient GEMMs:
    // when an SM frees up, a waiting recurrence block gets it before a GEMM tile does
    int prio_least = 0, prio_greatest = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    if (const char* ev = getenv("ARGSIM_WGRAD_OVERLAP")) wgrad_overlap = atoi(ev);
    if (const char* ev = getenv("ARGSIM_GROUP_CAP")) group_cap = atoi(ev);
    if (const char* ev = getenv("ARGSIM_SIDE_UNITS")) side_units = atoi(ev);
    if (const char* ev = getenv("ARGSIM_DEC_EARLY")) dec_early_on = atoi(ev);
    if (const char* ev = getenv("ARGSIM_ENC_BWD_CHUNK")) enc_bwd_chunk = atoi(ev);
    slice_budget = getenv("ARGSIM_NO_SLICE_BUDGET") == nullptr;
    early_adam = getenv("ARGSIM_NO_EARLY_ADAM") == nullptr;
    dp_one_allreduce = getenv("ARGSIM_DP_ONE_ALLREDUCE") != nullptr;
    seg_wgrad_on = getenv("ARGSIM_NO_SEG_WGRAD") == nullptr;
    if (const char* ev = getenv("ARGSIM_LOGIT_CHUNK")) logit_chunk = std::max(128, atoi(ev));
    const int prio_chain = wgrad_overlap ? prio_greatest : prio_least;
    // the NCCL stream sits in between: a bucket's all-reduce is on the way to the end of the step, the side stream is not
    const int prio_comm = wgrad_overlap ? (prio_least + prio_greatest) / 2 : prio_least;
    for (int i = 0; i < 3; ++i) CUDA_CHECK(cudaStreamCreateWithPriority(&st[i], cudaStreamNonBlocking, i == 2 ? prio_comm : prio_chain));
    for (auto& s : sw) CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio_chain));
    CUDA_CHECK(cudaStreamCreateWithPriority(&swg, cudaStreamNonBlocking, prio_least));
    CUDA_CHECK(cudaStreamCreateWithPriority(&sad, cudaStreamNonBlocking, prio_least));
    if (const char* ev = getenv("ARGSIM_DEC_SEG")) dec_seg = atoi(ev);
    if (const char* ev = getenv("ARGSIM_ENC_SEG")) enc_seg = atoi(ev);
    enc_seg_fwd = getenv("ARGSIM_ENC_SEG_FWD") != nullptr;
    pad_wave = getenv("ARGSIM_GRU_PAD_WAVE") ? atoi(getenv("ARGSIM_GRU_PAD_WAVE")) : 2;
    if (L > 8) dec_seg = 0;
    CUDA_CHECK(cudaEventCreateWithFlags(&ev_bucket, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&ev_comm, cudaEventDisableTiming));
    if (use_mma) mma = gru_mma_create(c.device);
    if (const char* ev = getenv("ARGSIM_GRU_TC")) gru_tc_mode = atoi(ev);
    if (use_mma && gru_tc_mode) tc = gru_tc_create(c.device);

    auto add = [&](const std::string& name, int64_t r, int64_t cdim) {
        ParamInfo pi;
        pi.name = name;
        pi.rank = cdim ? 2 : 1;
        pi.shape[0] = r; pi.shape[1] = cdim;
        pi.n = (size_t)r * (cdim ? cdim : 1);
        pi.off = nflat;
        nflat += align_up(pi.n, 64);
        pindex[name] = (int)params.size();
        params.push_back(pi);
    };
    if (!tied) {   // first in backward-completion order: part of the first all-reduce bucket
        add("logits/dense/kernel", D, V);
        add("logits/dense/bias", V, 0);
    }
    add("decode/out/kernel", D, D);
    add("decode/out/bias", D, 0);
    for (int j = L - 1; j >= 0; --j) {
        std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
        add(pre + "W", 3 * H, D); add(pre + "R", 3 * H, H); add(pre + "bW", 3 * H, 0); add(pre + "bR", 3 * H, 0);
    }
    add("latent/ex/kernel", R, D); add("latent/ex/bias", D, 0);
    add("latent/mu/kernel", EH, R); add("latent/mu/bias", R, 0);
    add("latent/lv/kernel", EH, R); add("latent/lv/bias", R, 0);
    if (attentive) {   // same all-reduce bucket as the latent affines
        if (EH % ATT_HEADS) throw std::runtime_error("attentive: the encoder width must be divisible by 8 heads");
        add("encode/cata/LayerNorm/gamma", EH, 0); add("encode/cata/LayerNorm/beta", EH, 0);
        for (const char* nm : {"p", "q", "k", "v"}) {
            add(std::string("encode/cata/") + nm + "/kernel", EH, EH);
            add(std::string("encode/cata/") + nm + "/bias", EH, 0);
        }
    }
    if (enc_kind != 0) {   // independent L-layer stack(s): layers in backward-completion order, directions adjacent
        for (int j = L - 1; j >= 0; --j)
            for (int d = 0; d < (enc_kind == 1 ? 2 : 1); ++d) {
                const std::string pre = enc_prefix(d, j);
                add(pre + "W", 3 * H, j == 0 ? D : H); add(pre + "R", 3 * H, H); add(pre + "bW", 3 * H, 0); add(pre + "bR", 3 * H, 0);
            }
    }
    for (int i = L; i >= 1 && enc_kind == 0; --i) {
        std::string pre = "encode/rnn" + std::to_string(i) + "/";
        const int in = (i == 1) ? D : 2 * H;
        // fwd and bwd tensors adjacent so that [W_f;W_b] is one (6H,in) GEMM operand
        add(pre + "fwd/W", 3 * H, in); add(pre + "bwd/W", 3 * H, in);
        add(pre + "fwd/R", 3 * H, H); add(pre + "bwd/R", 3 * H, H);
        add(pre + "fwd/bW", 3 * H, 0); add(pre + "bwd/bW", 3 * H, 0);
        add(pre + "fwd/bR", 3 * H, 0); add(pre + "bwd/bR", 3 * H, 0);
    }
    add("embed/embedding", V, D);
    // fwd/bwd tensor pairs are only contiguous when their sizes are multiples of the 64-element padding
    if ((3 * H) % 64) throw std::runtime_error("dim_emb must be a multiple of 64");
    CUDA_CHECK(cudaMalloc(&p, nflat * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&g, nflat * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&m, nflat * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&v, nflat * sizeof(float)));
    // on the engine's own stream: the work streams are cudaStreamNonBlocking and do not order against the legacy stream
    CUDA_CHECK(cudaMemsetAsync(p, 0, nflat * sizeof(float), st[0]));
    CUDA_CHECK(cudaMemsetAsync(g, 0, nflat * sizeof(float), st[0]));
    CUDA_CHECK(cudaMemsetAsync(m, 0, nflat * sizeof(float), st[0]));
    CUDA_CHECK(cudaMemsetAsync(v, 0, nflat * sizeof(float), st[0]));
    if (is_bf16) {
        CUDA_CHECK(cudaMalloc(&ph, nflat * sizeof(::bf16)));
        CUDA_CHECK(cudaMemsetAsync(ph, 0, nflat * sizeof(::bf16), st[0]));
    }
    CUDA_CHECK(cudaMalloc(&d_stats, 4 * sizeof(double)));
    CUDA_CHECK(cudaMallocHost(&h_stats, 4 * sizeof(double)));

    if (c.nranks > 1) {
        NcclApi& n = NcclApi::get();
        NcclApi::unique_id id;
        memcpy(id.internal, c.nccl_id, 128);
        NcclApi::comm_t comm;
        n.check(n.CommInitRank(&comm, c.nranks, id, c.rank), "ncclCommInitRank");
        nccl_comm = comm;
    }
    CUDA_CHECK(cudaDeviceSynchronize());
}

Engine::~Engine() {
    cudaDeviceSynchronize();
    if (nccl_comm) NcclApi::get().CommDestroy((NcclApi::comm_t)nccl_comm);
    if (mma) gru_mma_destroy(mma);
    if (tc) gru_tc_destroy(tc);
    cudaFree(p); cudaFree(g); cudaFree(m); cudaFree(v); cudaFree(ph);
    cudaFree(arena.base); cudaFree(d_stage); cudaFreeHost(h_stage); cudaFree(d_eps_in);
    for (auto& ps : pend) { if (ps.done) cudaEventDestroy(ps.done); if (ps.h_stats) cudaFreeHost(ps.h_stats); }
    cudaFree(d_stats); cudaFreeHost(h_stats); cudaFreeHost(h_out); cudaFree(gru_work);
    delete dbuf;
    for (auto& e : pev) cudaEventDestroy(e);
    for (auto& k : ktimers) { cudaEventDestroy(k.a); cudaEventDestroy(k.b); }
    cudaEventDestroy(ev_bucket); cudaEventDestroy(ev_comm);
    for (auto& e : ev_embed) if (e) cudaEventDestroy(e);
    for (auto& s : st) if (s) cudaStreamDestroy(s);
    for (auto& s : sw) if (s) cudaStreamDestroy(s);
    if (swg) cudaStreamDestroy(swg);
    if (sad) cudaStreamDestroy(sad);
    for (auto& e : evpool) cudaEventDestroy(e);
}

const ParamInfo& Engine::pinfo(const std::string& name) const {
    auto it = pindex.find(name);
    if (it == pindex.end()) throw std::runtime_error("unknown parameter: " + name);
    return params[it->second];
}
Mat Engine::pmat(const std::string& name) {
    const ParamInfo& pi = pinfo(name);
    const int cols = pi.rank == 2 ? (int)pi.shape[1] : (int)pi.shape[0];
    const long long rows = pi.rank == 2 ? pi.shape[0] : 1;
    return Mat(p + pi.off, ph ? ph + pi.off : nullptr, rows, cols, cols);
}
Mat Engine::gmat(const std::string& name) {
    const ParamInfo& pi = pinfo(name);
    const int cols = pi.rank == 2 ? (int)pi.shape[1] : (int)pi.shape[0];
    const long long rows = pi.rank == 2 ? pi.shape[0] : 1;
    return Mat(g + pi.off, nullptr, rows, cols, cols);
}

void Engine::refresh_shadow(size_t off, size_t n) {
    if (ph) launch_cast_bf16(p + off, ph + off, (long long)n, st[0]);
}

// A22: src/model.py:8-15,109-110.  Host RNG (mt19937_64); only the distributions are specified
// by the reference, TF's own streams cannot be matched.
void Engine::init_params(uint64_t seed_) {
    std::mt19937_64 rng(seed_);
    std::vector<float> host(nflat, 0.f);
    auto uni = [&](float* dst, size_t n, float bound) {
        std::uniform_real_distribution<float> d(-bound, bound);
        for (size_t i = 0; i < n; ++i) dst[i] = d(rng);
    };
    for (const ParamInfo& pi : params) {
        float* dst = host.data() + pi.off;
        if (pi.name == "encode/cata/LayerNorm/gamma") { for (size_t i = 0; i < pi.n; ++i) dst[i] = 1.f; continue; }   // layer_norm scale
        if (pi.rank == 1) continue;  // biases zero
        if (pi.name == "embed/embedding") {
            uni(dst, pi.n, sqrtf(6.0f / ((float)V / (float)D + 1.0f)));
        } else if (pi.name.size() > 2 && (pi.name.substr(pi.name.size() - 2) == "/W" || pi.name.substr(pi.name.size() - 2) == "/R")) {
            // variance_scaling(1,'fan_avg','uniform') per canonical (H,in) sub-matrix, one per gate
            const int in = (int)pi.shape[1];
            uni(dst, pi.n, sqrtf(6.0f / (float)(in + H)));
        } else {
            uni(dst, pi.n, sqrtf(6.0f / (float)(pi.shape[0] + pi.shape[1])));
        }
    }
    copy_sync(p, host.data(), nflat * sizeof(float), cudaMemcpyHostToDevice);
    CUDA_CHECK(cudaMemsetAsync(m, 0, nflat * sizeof(float), st[0]));
    CUDA_CHECK(cudaMemsetAsync(v, 0, nflat * sizeof(float), st[0]));
    refresh_shadow(0, nflat);
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    step = 0;
}

void Engine::set_param(const std::string& name, const float* src) {
    const ParamInfo& pi = pinfo(name);
    copy_sync(p + pi.off, src, pi.n * sizeof(float), cudaMemcpyHostToDevice);
    refresh_shadow(pi.off, pi.n);
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
}
void Engine::get_flat(const float* flat, const std::string& name, float* dst) {
    const ParamInfo& pi = pinfo(name);
    copy_sync(dst, flat + pi.off, pi.n * sizeof(float), cudaMemcpyDeviceToHost);
}
void Engine::set_flat(float* flat, const std::string& name, const float* src) {
    const ParamInfo& pi = pinfo(name);
    copy_sync(flat + pi.off, src, pi.n * sizeof(float), cudaMemcpyHostToDevice);
}

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
Mat Engine::act(long long rows, int cols) {
    if (use_tc) return Mat(nullptr, (::bf16*)arena.alloc(sizeof(::bf16) * rows * cols), rows, cols, cols);
    return Mat((float*)arena.alloc(sizeof(float) * rows * cols), nullptr, rows, cols, cols);
}
Mat Engine::f32(long long rows, int cols) {
    return Mat((float*)arena.alloc(sizeof(float) * rows * cols), nullptr, rows, cols, cols);
}
Mat Engine::both(long long rows, int cols) {
    float* f = (float*)arena.alloc(sizeof(float) * rows * cols);
    ::bf16* h = use_tc ? (::bf16*)arena.alloc(sizeof(::bf16) * rows * cols) : nullptr;
    return Mat(f, h, rows, cols, cols);
}

void Engine::gemm(const Mat& A, int a_mn, const Mat& B, int b_mn, const Mat& C, long long M, int N, long long K, float alpha,
                  const float* bias, int accumulate, cudaStream_t q) {
    if (arena.dry || M <= 0 || N <= 0) return;
    if (!q) q = st[0];
    if (use_tc) {
        if (!A.h || !B.h) throw std::runtime_error("gemm: bf16 operand view missing (internal error)");
        // per-group device timers (bench.py roofline): batched GEMMs outside the explicitly named groups, by role
        const bool timed = (cfg.flags & 8) && !in_ktimer && !decoding && M >= 256;   // decode steps may sit in a captured graph
        const bool on_side = q == swg;   // stretched by design: reported apart from the chain's GEMMs
        if (timed) kbegin(a_mn ? (on_side ? "k:gemm_wgrad_side" : "k:gemm_wgrad")
                          : b_mn ? (on_side ? "k:gemm_dgrad_or_dense_side" : "k:gemm_dgrad_or_dense")
                                 : (on_side ? "k:gemm_proj_side" : "k:gemm_proj"), q);
        gemm_tc(A.h, A.ld, a_mn, B.h, B.ld, b_mn, C.f, C.h, C.ld, (int)M, N, (int)K, alpha, bias, accumulate, q,
                q == swg ? side_units : 0);
        if (timed) kend(q, 2.0 * (double)M * N * K * 1e-9);
    } else {
        if (!A.f || !B.f) throw std::runtime_error("gemm: fp32 operand view missing (internal error)");
        gemm_simt(A.f, A.ld, a_mn, B.f, B.ld, b_mn, C.f, C.ld, (int)M, N, (int)K, alpha, bias, accumulate, C.h, q);
    }
}

// two long-lived events for the embedding pipeline (not part of the per-step pool that run_device() resets)
cudaEvent_t Engine::next_event_persistent(int i) {
    if (!ev_embed[i]) CUDA_CHECK(cudaEventCreateWithFlags(&ev_embed[i], cudaEventDisableTiming));
    return ev_embed[i];
}
cudaEvent_t Engine::next_event() {
    if (evcount == evpool.size()) {
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        evpool.push_back(e);
    }
    return evpool[evcount++];
}
// the decoder's layers run as a wavefront over time segments when the persistent kernel is in use
bool Engine::dec_wavefront(const SeqPlan& Dp) const {
    return use_mma && L > 1 && dec_seg > 0 && Dp.Tmax > dec_seg && gru_mma_fits(mma, 1, Dp.b);
}
// The stacked encoder's two directions as two chains of time-segment launches (one stream each) instead of one launch
// per layer: the forward direction has many live rows early and few late, the reverse direction the other way round, so
// per segment the 9 groups of 16 CTAs can be split to give BOTH directions slices of <= 8 rows (one MMA tile per step)
// almost everywhere, where a single launch is stuck with 4 slices of 16 rows per direction for all 512 steps.
bool Engine::enc_segmented(const SeqPlan& E) const {
    return use_mma && enc_kind == 0 && enc_seg > 0 && E.Tmax > enc_seg && gru_mma_fits(mma, 1, E.b);
}
// want8[d * nseg + i] for the i-th launch of chain d.  Forward pass: chain 0 walks segments upwards, chain 1 downwards;
// BPTT: the other way round.  Launch i of both chains run side by side.
void Engine::enc_slice_plan(const SeqPlan& E, int nseg, bool bptt, std::vector<int>* want8) const {
    want8->assign((size_t)2 * nseg, 0);
    // groups of 16 CTAs.  Round 1 gave data-parallel runs 8 so that 20 SMs stay free for the NCCL kernels of the overlapped
    // buckets; measured on 2 GPUs with the tail fixed (statistics / early Adam off the NCCL stream) 9 is faster: step 10.83 ->
    // 10.66 ms, encoder BPTT 4.25 -> 3.77 ms, exposed all-reduce wait 0.10 ms either way (ARGSIM_GROUP_CAP=8 restores it)
    const int cap = group_cap ? group_cap : 9;
    for (int i = 0; i < nseg; ++i) {
        int rows[2], g16[2], g8[2];
        for (int d = 0; d < 2; ++d) {
            const bool up = bptt ? (d == 1) : (d == 0);
            const int sg = up ? i : nseg - 1 - i;
            rows[d] = E.nact[sg * enc_seg];
            g16[d] = (rows[d] + 15) / 16; g8[d] = (rows[d] + 7) / 8;
        }
        const int big = rows[0] >= rows[1] ? 0 : 1, small = 1 - big;
        int use8[2] = {0, 0};
        if (g8[0] + g8[1] <= cap) use8[0] = use8[1] = 1;
        else if (g8[big] + g16[small] <= cap) use8[big] = 1;
        else if (g16[big] + g8[small] <= cap) use8[small] = 1;
        for (int d = 0; d < 2; ++d) (*want8)[(size_t)d * nseg + i] = use8[d] ? 1 : 2;   // 1 = 8-row slices, 2 = 16-row slices
    }
}

void Engine::colsum(const Mat& A, long long rows, int cols, float* out, int accumulate, cudaStream_t q) {
    if (arena.dry) return;
    if (!q) q = st[0];
    if (A.f) launch_colsum_f32(A.f, A.ld, rows, cols, out, q, accumulate);
    else launch_colsum_bf16(A.h, A.ld, rows, cols, out, q, accumulate);
}
void Engine::gather_embed(const int* ids, long long n, const Mat& out, cudaStream_t q) {
    if (arena.dry) return;
    if (!q) q = st[0];
    const ParamInfo& e = pinfo("embed/embedding");
    if (out.h) launch_embed_gather_bf16(ids, n, ph + e.off, D, out.h, q);
    else launch_embed_gather_f32(ids, n, p + e.off, D, out.f, q);
}
void Engine::rec_fwd(const GruFwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact, cudaStream_t q, int t0,
                     int Tseg, int slot, int want8, int pad) {
    if ((gru_tc_mode & 1) && tc) gru_tc_fwd(tc, dirs, ndir, P, d_off, d_nact, H, q, t0, Tseg, slot, 0, pad);
    else gru_mma_fwd(mma, dirs, ndir, P, d_off, d_nact, H, q, t0, Tseg, slot, want8, pad);
}
void Engine::gru_fwd(GruFwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact) {
    if (arena.dry) return;
    kbegin(ndir == 2 ? "k:gru_fwd_enc" : "k:gru_fwd_dec");
    const bool want_tc = use_mma && tc && ((gru_tc_mode & 1) || ((gru_tc_mode & 4) && gru_tc_throughput(tc, ndir, P.b) && !dirs[0].h0));
    if (want_tc && gru_tc_fits(tc, ndir, P.b)) gru_tc_fwd(tc, dirs, ndir, P, d_off, d_nact, H, st[0], 0, -1, 0, 0, /*pad: runs alone*/ 1);
    else if (use_mma && gru_mma_fits(mma, ndir, P.b)) gru_mma_fwd(mma, dirs, ndir, P, d_off, d_nact, H, st[0], 0, -1, 0, 0, /*pad: runs alone*/ 1);
    else gru_generic_fwd(dirs, ndir, P, H, gru_work, st);
    kend();
}
void Engine::rec_bwd(const GruBwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact, cudaStream_t q, int t0,
                     int Tseg, int slot, int want8, int pad, int chunk) {
    if ((gru_tc_mode & 2) && tc) gru_tc_bwd(tc, dirs, ndir, P, d_off, d_nact, H, q, t0, Tseg, slot, 0, pad);
    else gru_mma_bwd(mma, dirs, ndir, P, d_off, d_nact, H, q, t0, Tseg, slot, want8, pad, chunk);
}
void Engine::gru_bwd(GruBwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact) {
    if (arena.dry) return;
    kbegin(ndir == 2 ? "k:gru_bwd_enc" : "k:gru_bwd_dec");
    if (use_mma && (gru_tc_mode & 2) && tc && gru_tc_fits(tc, ndir, P.b)) rec_bwd(dirs, ndir, P, d_off, d_nact, st[0], 0, -1, 0, 0, /*pad: runs alone*/ 1, 0);
    else if (use_mma && gru_mma_fits(mma, ndir, P.b)) gru_mma_bwd(mma, dirs, ndir, P, d_off, d_nact, H, st[0], 0, -1, 0, 0, /*pad: runs alone*/ 1);
    else gru_generic_bwd(dirs, ndir, P, H, gru_work, st);
    kend();
}

void Engine::allreduce_bucket(size_t off0, size_t off1, cudaStream_t after) {
    if (arena.dry || cfg.nranks <= 1 || off1 <= off0) return;
    if (dp_one_allreduce) {   // A/B: no overlapped buckets, one all-reduce of every gradient behind the backward pass
        if (off1 != nflat) return;
        off0 = 0;
        after = nullptr;      // the main stream has joined the side stream by then
    }
    NcclApi& n = NcclApi::get();
    CUDA_CHECK(cudaEventRecord(ev_bucket, after ? after : st[0]));
    CUDA_CHECK(cudaStreamWaitEvent(st[2], ev_bucket, 0));
    n.check(n.AllReduce(g + off0, g + off0, off1 - off0, NcclApi::Float32, NcclApi::Sum, (NcclApi::comm_t)nccl_comm, st[2]),
            "ncclAllReduce");
}

void Engine::phase(const char* name) {
    if (arena.dry) return;
    if (pcount == pev.size()) {
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreate(&e));
        pev.push_back(e);
        pnames.push_back("");
    }
    pnames[pcount] = name;
    if (nvtx) nvtxMarkA((std::string("argsim:end:") + name).c_str());
    CUDA_CHECK(cudaEventRecord(pev[pcount], st[0]));
    ++pcount;
}
void Engine::kbegin(const char* name, cudaStream_t q) {
    if (!(cfg.flags & 8)) return;
    in_ktimer = true;
    if (kcount == ktimers.size()) {
        KTimer k;
        CUDA_CHECK(cudaEventCreate(&k.a));
        CUDA_CHECK(cudaEventCreate(&k.b));
        ktimers.push_back(k);
    }
    ktimers[kcount].name = name;
    ktimers[kcount].gflop = 0.0;
    CUDA_CHECK(cudaEventRecord(ktimers[kcount].a, q ? q : st[0]));
}
void Engine::kend(cudaStream_t q, double gflop) {
    if (!(cfg.flags & 8)) return;
    ktimers[kcount].gflop = gflop;
    CUDA_CHECK(cudaEventRecord(ktimers[kcount].b, q ? q : st[0]));
    ++kcount;
    in_ktimer = false;
}
void Engine::collect_timings() {
    tnames.clear(); tms.clear();
    for (size_t i = 0; i + 1 < pcount; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pev[i], pev[i + 1]);
        tnames.push_back(pnames[i + 1]);
        tms.push_back(ms);
    }
    std::map<std::string, std::pair<float, int>> agg;
    std::map<std::string, double> gf;
    for (size_t i = 0; i < kcount; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ktimers[i].a, ktimers[i].b);
        agg[ktimers[i].name].first += ms;
        agg[ktimers[i].name].second += 1;
        gf[ktimers[i].name] += ktimers[i].gflop;
    }
    for (auto& kv : agg) {
        tnames.push_back(kv.first);
        tms.push_back(kv.second.first);
        tnames.push_back(kv.first + "#n");
        tms.push_back((float)kv.second.second);
        if (gf[kv.first] > 0.0) {
            tnames.push_back(kv.first + "#gflop");
            tms.push_back((float)gf[kv.first]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// staging: host plan -> one pinned buffer -> one H2D copy
// ------------------------------------------------------------------------------------------
void Engine::stage(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, int need_dec, const DropoutSpec& drop,
                   const float* eps) {
    std::string e = build_batch_plan(src, b, Ts, tgt, Tt, cfg.bos, cfg.eos, need_dec, drop, &plan);
    if (!e.empty()) throw std::runtime_error(e);
    const SeqPlan& E = plan.enc;
    const SeqPlan& Dp = plan.dec;
    for (int x : plan.ids_src)
        if (x < 0 || x >= V) throw std::runtime_error("src token id out of range [0, dim_tgt)");
    for (int x : plan.ids_lead)
        if (x < 0 || x >= V) throw std::runtime_error("tgt token id out of range [0, dim_tgt)");
    for (int x : plan.labels)
        if (x < 0 || x >= V) throw std::runtime_error("eos/tgt token id out of range [0, dim_tgt)");
    size_t need = plan.ids_src.size() + plan.ids_lead.size() + plan.labels.size() + 3 * (size_t)b + (E.Tmax + 1) + E.Tmax +
                  (Dp.Tmax + 1) + Dp.Tmax + 64 * 10;
    if (need > stage_cap) {
        CUDA_CHECK(cudaDeviceSynchronize());
        cudaFreeHost(h_stage); cudaFree(d_stage);
        stage_cap = align_up(need * 2, 64);
        CUDA_CHECK(cudaMallocHost(&h_stage, 2 * stage_cap * sizeof(int)));
        CUDA_CHECK(cudaMalloc(&d_stage, 2 * stage_cap * sizeof(int)));
    }
    // slot of the step being staged: the previous step's tables stay in the other one while it runs
    const size_t slot_ = stage_slot >= 0 ? (size_t)stage_slot : (size_t)(submit_seq & 1);
    int* const hs = h_stage + slot_ * stage_cap;
    int* const ds = d_stage + slot_ * stage_cap;
    size_t o = 0;
    auto put = [&](const std::vector<int>& vsrc, int** dev) {
        *dev = ds + o;
        if (!vsrc.empty()) memcpy(hs + o, vsrc.data(), vsrc.size() * sizeof(int));
        o = align_up(o + vsrc.size(), 64);
    };
    put(plan.ids_src, &dp.ids_src);
    put(plan.ids_lead, &dp.ids_lead);
    put(plan.labels, &dp.labels);
    put(plan.enc_last, &dp.enc_last);
    put(Dp.perm, &dp.dec_perm);
    put(E.off, &dp.enc_off);
    put(E.nact, &dp.enc_nact);
    put(Dp.off, &dp.dec_off);
    put(Dp.nact, &dp.dec_nact);
    {   // global row index per row: keys the eps stream on the device (the keep mask was drawn above, on the host)
        std::vector<int> rid(b);
        for (int i = 0; i < b; ++i) rid[i] = (int)(drop.rows ? drop.rows[i] : drop.row0 + i);
        put(rid, &dp.row_ids);
    }
    CUDA_CHECK(cudaMemcpyAsync(ds, hs, o * sizeof(int), cudaMemcpyHostToDevice, st[0]));
    have_eps = false;
    if (eps) {
        size_t n = (size_t)b * R;
        if (n > eps_cap) {
            cudaFree(d_eps_in);
            CUDA_CHECK(cudaMalloc(&d_eps_in, n * sizeof(float)));
            eps_cap = n;
        }
        CUDA_CHECK(cudaMemcpyAsync(d_eps_in, eps, n * sizeof(float), cudaMemcpyHostToDevice, st[0]));
        have_eps = true;
    }
    size_t gw = gru_generic_work_floats(b, H) * 2;
    if (gw > gru_work_cap) {
        CUDA_CHECK(cudaStreamSynchronize(st[0]));
        cudaFree(gru_work);
        CUDA_CHECK(cudaMalloc(&gru_work, gw * sizeof(float)));
        gru_work_cap = gw;
    }
}

void Engine::ensure_arena(int mode) {
    arena.reset(true);
    arena.high = 0;
    program(mode, false);
    size_t need = arena.high + 4096;
    if (need > arena.cap) {
        CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(arena.base);
        arena.cap = need + need / 8;
        CUDA_CHECK(cudaMalloc(&arena.base, arena.cap));
    }
    arena.reset(false);
}

void Engine::run_device(int mode, bool apply_update) {
    ensure_arena(mode);
    pcount = 0; kcount = 0; evcount = 0;
    if (nvtx) nvtxRangePushA(mode == 2 ? "argsim:train_step" : mode == 1 ? "argsim:eval_step" : "argsim:embed");
    try {
        program(mode, apply_update);
    } catch (...) {
        if (nvtx) nvtxRangePop();
        throw;
    }
    if (nvtx) nvtxRangePop();
}

// ------------------------------------------------------------------------------------------
// the device program
// ------------------------------------------------------------------------------------------
void Engine::program(int mode, bool apply_update) {
    const bool train = (mode == 2);
    const SeqPlan& E = plan.enc;
    const SeqPlan& Dp = plan.dec;
    const long long S = E.rows, N = Dp.rows;
    const int b = plan.b;
    cudaStream_t s = st[0];
    const int64_t b_glob = last.b_global > 0 ? last.b_global : b;
    const int64_t n_glob = last.n_tok_global > 0 ? last.n_tok_global : N;
    float keepwd, anneal, lr;
    schedule_f32(step, cfg.accelerate, cfg.learn_rate, &keepwd, &anneal, &lr);

    phase("begin");
    RUN(CUDA_CHECK(cudaMemsetAsync(d_stats, 0, 4 * sizeof(double), s)));
    if (train) RUN(CUDA_CHECK(cudaMemsetAsync(g, 0, nflat * sizeof(float), s)));
    bucket_lo = 0;

    // The decoder's first-layer inputs do not depend on the encoder: emb_tgt is gathered and projected on the side
    // stream while the encoder's recurrences run (they are off the serial chain that way)
    const bool dec_early = mode != 0 && N > 0 && wgrad_overlap && use_mma && dec_wavefront(Dp) && dec_early_on;
    Mat decY0, decGX0;
    cudaEvent_t ev_dec0 = nullptr;
    if (dec_early) {
        decY0 = act(N, D);
        decGX0 = f32(N, 3 * H);
        if (!arena.dry) {
            cudaEvent_t ev = next_event();
            CUDA_CHECK(cudaEventRecord(ev, s));
            CUDA_CHECK(cudaStreamWaitEvent(swg, ev, 0));
            gather_embed(dp.ids_lead, N, decY0, swg);
            gemm(decY0, 0, pmat("decode/rnn/l0/W"), 0, decGX0, N, 3 * H, D, 1.f, p + pinfo("decode/rnn/l0/bW").off, 0, swg);
            ev_dec0 = next_event();
            CUDA_CHECK(cudaEventRecord(ev_dec0, swg));
        }
    }

    // ---------------- encoder (model.py:111-122): 3 x (fwd GRU || bwd GRU) over packed rows
    std::vector<Mat> encX(L + 1);
    std::vector<float*> encCache(2 * L, nullptr);
    encX[0] = act(S, D);
    gather_embed(dp.ids_src, S, encX[0]);
    const int nd_enc = (enc_kind == 2) ? 1 : 2;
    std::vector<Mat> encIn[2];   // branches 1 / 2: encIn[d][j] = input of layer j of stack d, [L] = its top output
    if (enc_kind != 0) {
        // model.py:124-131: one or two independent L-layer stacks over the packed rows (the 'bwd' stack simply walks t
        // downwards), concatenated only at the top; the two stacks of a layer run as the two directions of one launch
        Mat HS = act(S, EH);
        for (int d = 0; d < nd_enc; ++d) { encIn[d].resize(L + 1); encIn[d][0] = encX[0]; }
        for (int j = 0; j < L; ++j) {
            GruFwdArgs a[2];
            for (int d = 0; d < nd_enc; ++d) {
                const std::string pre = enc_prefix(d, j);
                Mat GX = f32(S, 3 * H);
                gemm(encIn[d][j], 0, pmat(pre + "W"), 0, GX, S, 3 * H, j == 0 ? D : H, 1.f, p + pinfo(pre + "bW").off, 0);
                Mat out = (j == L - 1) ? HS.colslice(d * H, H) : act(S, H);
                if (train) encCache[d * L + j] = (float*)arena.alloc(sizeof(float) * S * 4 * H);
                a[d].gx = GX.f; a[d].ld_gx = 3 * H;
                a[d].R_f = p + pinfo(pre + "R").off;
                a[d].R_h = ph ? ph + pinfo(pre + "R").off : nullptr;
                a[d].bR = p + pinfo(pre + "bR").off;
                a[d].h0 = nullptr;
                a[d].hs_f = out.f; a[d].hs_h = out.h; a[d].ld_hs = out.ld;
                a[d].cache = encCache[d * L + j];
                a[d].reverse = (enc_kind == 1 && d == 1) ? 1 : 0;
                encIn[d][j + 1] = out;
            }
            gru_fwd(a, nd_enc, E, dp.enc_off, dp.enc_nact);
        }
        encX[L] = HS;
    }
    for (int i = 0; i < L && enc_kind == 0; ++i) {
        const int in = (i == 0) ? D : 2 * H;
        const std::string pre = "encode/rnn" + std::to_string(i + 1) + "/";
        Mat W = pmat(pre + "fwd/W");
        W.rows = 6 * H;
        Mat GX = f32(S, 6 * H);
        gemm(encX[i], 0, W, 0, GX, S, 6 * H, in, 1.f, p + pinfo(pre + "fwd/bW").off, 0);
        Mat HS = act(S, 2 * H);
        GruFwdArgs a[2];
        for (int d = 0; d < 2; ++d) {
            const std::string pd = pre + (d ? "bwd/" : "fwd/");
            if (train) encCache[2 * i + d] = (float*)arena.alloc(sizeof(float) * S * 4 * H);
            a[d].gx = GX.f + d * 3 * H; a[d].ld_gx = 6 * H;
            a[d].R_f = p + pinfo(pd + "R").off;
            a[d].R_h = ph ? ph + pinfo(pd + "R").off : nullptr;
            a[d].bR = p + pinfo(pd + "bR").off;
            a[d].h0 = nullptr;
            a[d].hs_f = HS.f ? HS.f + d * H : nullptr;
            a[d].hs_h = HS.h ? HS.h + d * H : nullptr;
            a[d].ld_hs = 2 * H;
            a[d].cache = encCache[2 * i + d];
            a[d].reverse = d;
        }
        if (enc_seg_fwd && enc_segmented(E)) {
            const int nsegE = (E.Tmax + enc_seg - 1) / enc_seg;
            float* hTe[2][2];
            for (int d = 0; d < 2; ++d)
                for (int q = 0; q < 2; ++q) hTe[d][q] = (float*)arena.alloc(sizeof(float) * b * H);
            if (!arena.dry) {
                std::vector<int> want8;
                enc_slice_plan(E, nsegE, false, &want8);
                for (int d = 0; d < 2; ++d)
                    for (int q = 0; q < 2; ++q) CUDA_CHECK(cudaMemsetAsync(hTe[d][q], 0, sizeof(float) * b * H, s));
                kbegin("k:gru_fwd_enc");
                cudaEvent_t fork = next_event();
                CUDA_CHECK(cudaEventRecord(fork, s));
                for (int d = 0; d < 2; ++d) CUDA_CHECK(cudaStreamWaitEvent(sw[d], fork, 0));
                for (int k = 0; k < nsegE; ++k)
                    for (int d = 0; d < 2; ++d) {
                        const int sg = d ? nsegE - 1 - k : k;   // the reverse direction starts at the end of the sequence
                        const int t0 = sg * enc_seg, tl = std::min(enc_seg, E.Tmax - t0);
                        GruFwdArgs x = a[d];
                        x.h0 = (k > 0) ? hTe[d][(k - 1) & 1] : nullptr;
                        x.hT = (k + 1 < nsegE) ? hTe[d][k & 1] : nullptr;
                        rec_fwd(&x, 1, E, dp.enc_off, dp.enc_nact, sw[d], t0, tl, d, want8[(size_t)d * nsegE + k], pad_wave);
                    }
                for (int d = 0; d < 2; ++d) {
                    cudaEvent_t ev = next_event();
                    CUDA_CHECK(cudaEventRecord(ev, sw[d]));
                    CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
                }
                kend();
            }
        } else {
            gru_fwd(a, 2, E, dp.enc_off, dp.enc_nact);
        }
        encX[i + 1] = HS;
    }
    phase("enc_fwd");

    // ---------------- final state + latent (model.py:133-156)
    Mat henc = both(b, EH);
    RUN(launch_row_gather(encX[L].f, encX[L].h, EH, dp.enc_last, henc.f, henc.h, EH, nullptr, b, EH, s));
    // attentive (model.py:136-145): h <- layer_norm(h + p(attend(q(h), k(hs), v(hs)))) over the sequence's own steps
    struct { Mat h0, q, K, Vv, y; float *prob = nullptr, *xhat = nullptr, *rstd = nullptr; } att;
    auto cata = [&](const char* nm, const char* what) { return std::string("encode/cata/") + nm + "/" + what; };
    if (attentive) {
        const Mat& HSx = encX[L];
        att.h0 = henc;
        att.q = f32(b, EH); att.K = f32(S, EH); att.Vv = f32(S, EH); att.y = both(b, EH);
        att.prob = (float*)arena.alloc(sizeof(float) * S * ATT_HEADS);
        att.xhat = (float*)arena.alloc(sizeof(float) * b * EH);
        att.rstd = (float*)arena.alloc(sizeof(float) * b);
        gemm(henc, 0, pmat(cata("q", "kernel")), 1, att.q, b, EH, EH, 1.f, p + pinfo(cata("q", "bias")).off, 0);
        gemm(HSx, 0, pmat(cata("k", "kernel")), 1, att.K, S, EH, EH, 1.f, p + pinfo(cata("k", "bias")).off, 0);
        gemm(HSx, 0, pmat(cata("v", "kernel")), 1, att.Vv, S, EH, EH, 1.f, p + pinfo(cata("v", "bias")).off, 0);
        RUN(launch_attn_fwd(att.q.f, att.K.f, att.Vv.f, b, EH, ATT_HEADS, dp.enc_off, E.Tmax, dp.enc_last, att.prob, att.y.f, att.y.h, s));
        Mat pp = f32(b, EH), hn = both(b, EH);
        gemm(att.y, 0, pmat(cata("p", "kernel")), 1, pp, b, EH, EH, 1.f, p + pinfo(cata("p", "bias")).off, 0);
        RUN(launch_resid_ln_fwd(henc.f, pp.f, p + pinfo("encode/cata/LayerNorm/gamma").off, p + pinfo("encode/cata/LayerNorm/beta").off,
                                b, EH, att.xhat, att.rstd, hn.f, hn.h, s));
        henc = hn;
    }
    Mat mulv = f32(b, 2 * R);
    gemm(henc, 0, pmat("latent/mu/kernel"), 1, mulv.colslice(0, R), b, R, EH, 1.f, p + pinfo("latent/mu/bias").off, 0);
    outp = Out();
    outp.mulv = mulv.f;
    if (mode == 0) {
        phase("latent_fwd");
        return;
    }
    gemm(henc, 0, pmat("latent/lv/kernel"), 1, mulv.colslice(R, R), b, R, EH, 1.f, p + pinfo("latent/lv/bias").off, 0);
    Mat z = both(b, R);
    float* eps_used = train ? (float*)arena.alloc(sizeof(float) * b * R) : nullptr;
    float* kld_samp = (float*)arena.alloc(sizeof(float) * b * R);
    outp.kld_samp = kld_samp;
    RUN(launch_latent_fwd(mulv.f, have_eps ? d_eps_in : nullptr, b, R, train ? 1 : 0, seed, (uint64_t)step, last.row0, dp.row_ids, eps_used,
                          z.f, z.h, kld_samp, d_stats, s));
    Mat hx = f32(b, D);
    gemm(z, 0, pmat("latent/ex/kernel"), 1, hx, b, D, R, 1.f, p + pinfo("latent/ex/bias").off, 0);
    Mat hx_sorted = f32(b, H);
    RUN(launch_row_gather(hx.f, nullptr, D, dp.dec_perm, hx_sorted.f, nullptr, H, nullptr, b, H, s));
    phase("latent_fwd");

    // ---------------- decoder (model.py:158-162): 3 stacked GRUs, all seeded with ex(z)
    std::vector<Mat> decY(L + 1);
    std::vector<float*> decCache(L, nullptr);
    if (dec_early) {
        decY[0] = decY0;
        if (ev_dec0) CUDA_CHECK(cudaStreamWaitEvent(s, ev_dec0, 0));
    } else {
        decY[0] = act(N, D);
        gather_embed(dp.ids_lead, N, decY[0]);
    }
    std::vector<Mat> decGX(L);
    for (int j = 0; j < L; ++j) {
        decGX[j] = (j == 0 && dec_early) ? decGX0 : f32(N, 3 * H);
        decY[j + 1] = act(N, H);
        if (train) decCache[j] = (float*)arena.alloc(sizeof(float) * N * 4 * H);
    }
    auto dec_fwd_args = [&](int j) {
        const std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
        GruFwdArgs a;
        a.gx = decGX[j].f; a.ld_gx = 3 * H;
        a.R_f = p + pinfo(pre + "R").off;
        a.R_h = ph ? ph + pinfo(pre + "R").off : nullptr;
        a.bR = p + pinfo(pre + "bR").off;
        a.h0 = hx_sorted.f;
        a.hs_f = decY[j + 1].f; a.hs_h = decY[j + 1].h; a.ld_hs = H;
        a.cache = decCache[j];
        a.reverse = 0;
        return a;
    };
    const bool wave = dec_wavefront(Dp);
    // a trailing segment shorter than a quarter of the others is merged into its neighbour (IAC batches: 513 = 8 * 64 + 1
    // steps would otherwise pay a whole pipeline stage of launches for one step)
    int nseg = wave ? (Dp.Tmax + dec_seg - 1) / dec_seg : 1;
    if (wave && nseg > 1 && (Dp.Tmax - (nseg - 1) * dec_seg) * 4 < dec_seg) --nseg;
    auto seg_len = [&](int sg) { return sg + 1 < nseg ? dec_seg : Dp.Tmax - sg * dec_seg; };
    // Slices per wavefront launch.  A slice of <= 8 live rows needs one n=8 MMA tile per step (~2,700 cycles), a slice of
    // 9..16 rows two (~4,700), but 8-row slices cost twice the CTAs.  Launches (layer j, segment sg) with the same
    // j + sg run side by side: within the 9 groups of 16 CTAs the chip holds, the launches with the most live rows get
    // 8-row slices first.  want8[j * nseg + sg] = 1 -> 8 rows per slice.
    std::vector<int> want8((size_t)L * nseg, 0);
    if (wave && slice_budget) {
        const int max_groups = group_cap ? group_cap : 9;
        for (int stage = 0; stage < nseg + L - 1; ++stage) {
            std::vector<std::pair<int, int>> items;   // (live rows, j)
            int total = 0;
            for (int j = 0; j < L; ++j) {
                const int sg = stage - j;
                if (sg < 0 || sg >= nseg) continue;
                const int rows = Dp.nact[sg * dec_seg];
                const bool small = rows <= 16;
                want8[(size_t)j * nseg + sg] = small ? 1 : 2;   // 1 = 8-row slices, 2 = 16-row slices
                total += small ? (rows + 7) / 8 : (rows + 15) / 16;
                if (!small) items.push_back({rows, j});
            }
            std::sort(items.begin(), items.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first > y.first; });
            for (auto& it : items) {
                const int g16 = (it.first + 15) / 16, g8 = (it.first + 7) / 8;
                if (total - g16 + g8 <= max_groups) {
                    total += g8 - g16;
                    want8[(size_t)it.second * nseg + (stage - it.second)] = 1;
                }
            }
        }
    }
    if (!wave) {
        for (int j = 0; j < L; ++j) {
            const std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
            gemm(decY[j], 0, pmat(pre + "W"), 0, decGX[j], N, 3 * H, D, 1.f, p + pinfo(pre + "bW").off, 0);
            GruFwdArgs a = dec_fwd_args(j);
            gru_fwd(&a, 1, Dp, dp.dec_off, dp.dec_nact);
        }
    } else {
        // Wavefront: layer l works on time segment s while layer l-1 already runs segment s+1 (one stream per
        // layer, events between them).  The chain shrinks from L*T to about T + (L-1)*segment serial steps.
        float* hT[8][2];
        for (int j = 0; j < L; ++j)
            for (int q = 0; q < 2; ++q) hT[j][q] = (float*)arena.alloc(sizeof(float) * b * H);
        if (!dec_early) gemm(decY[0], 0, pmat("decode/rnn/l0/W"), 0, decGX[0], N, 3 * H, D, 1.f, p + pinfo("decode/rnn/l0/bW").off, 0);
        if (!arena.dry) {
            kbegin("k:gru_fwd_dec");
            cudaEvent_t fork = next_event();
            CUDA_CHECK(cudaEventRecord(fork, s));
            for (int j = 0; j < L; ++j) CUDA_CHECK(cudaStreamWaitEvent(sw[j], fork, 0));
            std::vector<cudaEvent_t> done(L * nseg);
            for (int sg = 0; sg < nseg; ++sg) {
                const int t0 = sg * dec_seg, tl = seg_len(sg);
                const long long r0 = Dp.off[t0], nr = Dp.off[t0 + tl] - r0;
                for (int j = 0; j < L; ++j) {
                    cudaStream_t q = sw[j];
                    const std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
                    if (j > 0) {
                        CUDA_CHECK(cudaStreamWaitEvent(q, done[(j - 1) * nseg + sg], 0));
                        gemm(decY[j].rowslice(r0, nr), 0, pmat(pre + "W"), 0, decGX[j].rowslice(r0, nr), nr, 3 * H, D, 1.f,
                             p + pinfo(pre + "bW").off, 0, q);
                    }
                    GruFwdArgs a = dec_fwd_args(j);
                    if (sg > 0) a.h0 = hT[j][(sg - 1) & 1];
                    a.hT = (sg + 1 < nseg) ? hT[j][sg & 1] : nullptr;
                    rec_fwd(&a, 1, Dp, dp.dec_off, dp.dec_nact, q, t0, tl, j, want8[(size_t)j * nseg + sg], pad_wave);
                    done[j * nseg + sg] = next_event();
                    CUDA_CHECK(cudaEventRecord(done[j * nseg + sg], q));
                }
            }
            for (int j = 0; j < L; ++j) CUDA_CHECK(cudaStreamWaitEvent(s, done[j * nseg + nseg - 1], 0));
            kend();
        }
    }
    Mat HO = act(N, D);
    gemm(decY[L], 0, pmat("decode/out/kernel"), 1, HO, N, D, D, 1.f, p + pinfo("decode/out/bias").off, 0);
    phase("dec_fwd");

    // ---------------- vocab projection + fused softmax-CE (+ its two backward GEMMs), row chunks
    // sized so that the logits chunk stays L2 resident (126 MB) between its producer and consumers.
    float* loss_samp = (float*)arena.alloc(sizeof(float) * N);
    float* err_samp = (float*)arena.alloc(sizeof(float) * N);
    int* pred = (int*)arena.alloc(sizeof(int) * N);
    outp.loss_samp = loss_samp; outp.err_samp = err_samp; outp.pred = pred;
    const float scale = 1.0f / sqrtf((float)D);
    long long chunk = use_tc ? 4096 : 2048;
    if (logit_chunk) chunk = logit_chunk;
    else if (use_tc && N > 0) {
        // equal chunks of at most 4608 rows (75 MB of bf16 logits: L2 resident) instead of 4096-row chunks plus a tail: the C1
        // batch (N = 8,872) ran 4096 + 4096 + 680 rows, and the 680-row launches of the GEMMs and of the fused softmax-CE
        // (less than one wave of 5 rows per SM) dragged the group averages down (softmax-CE 0.72 of HBM peak in-step)
        // (capping the chunk so that its logits stay L2 resident at V = 32768 -- 1,152 rows -- was measured slower on the scaled
        // config: logits phase 11.6 -> 13.0 ms, 29 launches per GEMM instead of 8)
        const long long nch = (N + 4607) / 4608;
        chunk = ((N + nch - 1) / nch + 127) / 128 * 128;
    }
    chunk = std::min<long long>(chunk, std::max<long long>(N, 1));
    // Weight gradients are not on the serial chain: with the persistent recurrence in use they go to the low-priority
    // side stream and fill the SMs the recurrence launches of the layers below leave free (a third of the encoder's
    // BPTT runs on 4 of its 9 groups); the all-reduce bucket of a layer is then ordered after the side stream.
    const bool side_on = wgrad_overlap && use_mma && train;
    const bool side = side_on && !arena.dry;
    auto side_after_main = [&]() {
        cudaEvent_t ev = next_event();
        CUDA_CHECK(cudaEventRecord(ev, s));
        CUDA_CHECK(cudaStreamWaitEvent(swg, ev, 0));
    };
    // with the side stream every chunk keeps its own d logits until its weight-gradient GEMM has read it (<= 4 GB)
    const long long nchunk = (N + chunk - 1) / chunk;
    const bool side_logits = side_on && (double)nchunk * chunk * V * 2 <= 4e9;
    std::vector<Mat> lbuf(side_logits ? nchunk : 1);
    for (auto& m : lbuf) m = act(chunk, V);
    Mat dHO = train ? act(N, D) : Mat();
    Mat Emb = pmat("embed/embedding");
    Mat gE = gmat("embed/embedding");
    Mat Kl = tied ? Mat() : pmat("logits/dense/kernel");
    Mat gKl = tied ? Mat() : gmat("logits/dense/kernel");
    auto logits_wgrad = [&](const Mat& dlogits, long long r0, long long nr, cudaStream_t qw) {
        RUN(kbegin(qw ? "k:logits_wgrad_side" : "k:logits_wgrad", qw));
        if (tied) {
            gemm(dlogits, 1, HO.rowslice(r0, nr), 1, gE, V, D, nr, scale, nullptr, 1, qw);
        } else {   // dK += ho^T . dlogits ; db += column sums of dlogits
            gemm(HO.rowslice(r0, nr), 1, dlogits, 1, gKl, D, V, nr, 1.f, nullptr, 1, qw);
            colsum(dlogits, nr, V, gptr("logits/dense/bias"), 1, qw);
        }
        RUN(kend(qw));
    };
    for (long long r0 = 0; r0 < N; r0 += chunk) {
        const long long nr = std::min(chunk, N - r0);
        const Mat& logits = lbuf[side_logits ? r0 / chunk : 0];
        cudaStream_t qw = (side && side_logits) ? swg : nullptr;
        RUN(kbegin("k:logits_gemm"));
        if (tied) gemm(HO.rowslice(r0, nr), 0, Emb, 0, logits, nr, V, D, scale, nullptr, 0);
        else gemm(HO.rowslice(r0, nr), 0, Kl, 1, logits, nr, V, D, 1.f, p + pinfo("logits/dense/bias").off, 0);   // h.K + b, K is (in,out)
        RUN(kend());
        RUN(kbegin("k:softmax_ce"));
        if (logits.h)
            RUN(launch_ce_bf16(logits.h, V, dp.labels + r0, nr, V, 1.0f / (float)n_glob, train, loss_samp + r0, err_samp + r0,
                               pred + r0, d_stats, s));
        else
            RUN(launch_ce_f32(logits.f, V, dp.labels + r0, nr, V, 1.0f / (float)n_glob, train, loss_samp + r0, err_samp + r0,
                              pred + r0, d_stats, s));
        RUN(kend());
        if (train) {
            // dE (dense part) += D^-1/2 * dlogits^T . ho ;  dho = D^-1/2 * dlogits . E
            if (!qw) logits_wgrad(logits, r0, nr, nullptr);
            RUN(kbegin("k:logits_dgrad"));
            if (tied) gemm(logits, 0, Emb, 1, dHO.rowslice(r0, nr), nr, D, V, scale, nullptr, 0);
            else gemm(logits, 0, Kl, 0, dHO.rowslice(r0, nr), nr, D, V, 1.f, nullptr, 0);   // dho = dlogits . K^T
            RUN(kend());
        }
    }
    if (side && side_logits) {
        // the vocabulary weight gradient of every chunk follows once the chain has left the logits phase: issued chunk by
        // chunk it would share the SMs with the chain's own GEMMs (vocab GEMM 0.60 -> 0.54 of peak, dgrad 0.50 -> 0.34)
        side_after_main();
        for (long long r0 = 0; r0 < N; r0 += chunk) logits_wgrad(lbuf[r0 / chunk], r0, std::min(chunk, N - r0), swg);
    }
    phase("logits_ce");
    if (!train) return;
    if (cfg.nranks > 1 && !arena.dry) {
        // the step statistics (loss / error / KL sums) are final once the forward pass is: their 32-byte all-reduce goes
        // out now, not at the end of the step where it would sit behind the last gradient bucket on the tail
        NcclApi& n = NcclApi::get();
        cudaEvent_t ev = next_event();
        CUDA_CHECK(cudaEventRecord(ev, s));
        CUDA_CHECK(cudaStreamWaitEvent(st[2], ev, 0));
        n.check(n.AllReduce(d_stats, d_stats, 4, NcclApi::Float64, NcclApi::Sum, (NcclApi::comm_t)nccl_comm, st[2]),
                "ncclAllReduce(stats)");
    }

    // ---------------- backward: out affine
    Mat dY = f32(N, D);   // H == D (model.py:160: the decoder GRUs are dim_emb wide)
    if (side) {
        side_after_main();
        gemm(dHO, 0, pmat("decode/out/kernel"), 0, dY, N, D, D, 1.f, nullptr, 0);
    }
    gemm(decY[L], 1, dHO, 1, gmat("decode/out/kernel"), D, D, N, 1.f, nullptr, 1, side ? swg : nullptr);
    colsum(dHO, N, D, gptr("decode/out/bias"), 0, side ? swg : nullptr);
    if (!side) gemm(dHO, 0, pmat("decode/out/kernel"), 0, dY, N, D, D, 1.f, nullptr, 0);
    allreduce_bucket(bucket_lo, pinfo("decode/out/bias").off + align_up(D, 64), side ? swg : nullptr);
    bucket_lo = pinfo("decode/out/bias").off + align_up(D, 64);

    // ---------------- backward: decoder GRUs (BPTT), dh0 of all layers sums into d ex(z)
    Mat dhx_sorted = f32(b, H);
    RUN(CUDA_CHECK(cudaMemsetAsync(dhx_sorted.f, 0, sizeof(float) * b * H, s)));
    Mat dYn = f32(N, D);
    auto dec_bwd_args = [&](int j, const Mat& dhs, const Mat& dGX, const Mat& dGH, const Mat& HP, float* dh0) {
        const std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
        GruBwdArgs a;
        a.dhs = dhs.f; a.ld_dhs = dhs.ld;
        a.hs_f = decY[j + 1].f; a.hs_h = decY[j + 1].h; a.ld_hs = H;
        a.h0 = hx_sorted.f;
        a.cache = decCache[j];
        a.R_f = p + pinfo(pre + "R").off;
        a.R_h = ph ? ph + pinfo(pre + "R").off : nullptr;
        a.dgx_f = dGX.f; a.dgx_h = dGX.h; a.dgh_f = dGH.f; a.dgh_h = dGH.h; a.ld_dg = 3 * H;
        a.hp_f = HP.f; a.hp_h = HP.h; a.ld_hp = H;
        a.dh0 = dh0;
        a.reverse = 0;
        return a;
    };
    auto dec_wgrad = [&](int j, const Mat& dGX, const Mat& dGH, const Mat& HP, cudaStream_t q = nullptr) {
        const std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
        gemm(dGX, 1, decY[j], 1, gmat(pre + "W"), 3 * H, D, N, 1.f, nullptr, 1, q);
        gemm(dGH, 1, HP, 1, gmat(pre + "R"), 3 * H, H, N, 1.f, nullptr, 1, q);
        colsum(dGX, N, 3 * H, gptr(pre + "bW"), 0, q);
        colsum(dGH, N, 3 * H, gptr(pre + "bR"), 0, q);
    };

    if (!wave) {
        Mat dGX = act(N, 3 * H), dGH = act(N, 3 * H), HP = act(N, H);
        for (int j = L - 1; j >= 0; --j) {
            const std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
            GruBwdArgs a = dec_bwd_args(j, dY, dGX, dGH, HP, dhx_sorted.f);
            gru_bwd(&a, 1, Dp, dp.dec_off, dp.dec_nact);
            dec_wgrad(j, dGX, dGH, HP);
            gemm(dGX, 0, pmat(pre + "W"), 1, dYn, N, D, 3 * H, 1.f, nullptr, 0);
            std::swap(dY, dYn);
            const size_t end = pinfo(pre + "bR").off + align_up(3 * H, 64);
            allreduce_bucket(bucket_lo, end);
            bucket_lo = end;
        }
    } else {
        // Wavefront in reverse: layer l's BPTT over segment s starts as soon as layer l+1 has produced the
        // gradient of its inputs for that segment (dgrad GEMM per segment on layer l's stream).
        std::vector<Mat> dGXl(L), dGHl(L), HPl(L), dYl(L);
        float *dh0l[8], *carry[8][2];
        for (int j = 0; j < L; ++j) {
            dGXl[j] = act(N, 3 * H); dGHl[j] = act(N, 3 * H); HPl[j] = act(N, H);
            dYl[j] = (j == L - 1) ? dY : f32(N, D);
            dh0l[j] = (float*)arena.alloc(sizeof(float) * b * H * 3);
            carry[j][0] = dh0l[j] + (size_t)b * H;
            carry[j][1] = dh0l[j] + (size_t)2 * b * H;
            RUN(CUDA_CHECK(cudaMemsetAsync(dh0l[j], 0, sizeof(float) * b * H * 3, s)));
        }
        if (!arena.dry) {
            kbegin("k:gru_bwd_dec");
            cudaEvent_t fork = next_event();
            CUDA_CHECK(cudaEventRecord(fork, s));
            for (int j = 0; j < L; ++j) CUDA_CHECK(cudaStreamWaitEvent(sw[j], fork, 0));
            std::vector<cudaEvent_t> done(L * nseg);
            for (int sg = nseg - 1; sg >= 0; --sg) {
                const int t0 = sg * dec_seg, tl = seg_len(sg);
                const long long r0 = Dp.off[t0], nr = Dp.off[t0 + tl] - r0;
                for (int j = L - 1; j >= 0; --j) {
                    cudaStream_t q = sw[j];
                    if (j < L - 1) {
                        const std::string up = "decode/rnn/l" + std::to_string(j + 1) + "/";
                        CUDA_CHECK(cudaStreamWaitEvent(q, done[(j + 1) * nseg + sg], 0));
                        gemm(dGXl[j + 1].rowslice(r0, nr), 0, pmat(up + "W"), 1, dYl[j].rowslice(r0, nr), nr, D, 3 * H, 1.f, nullptr, 0, q);
                    }
                    GruBwdArgs a = dec_bwd_args(j, dYl[j], dGXl[j], dGHl[j], HPl[j], dh0l[j]);
                    a.dh_in = (sg + 1 < nseg) ? carry[j][(sg + 1) & 1] : nullptr;
                    a.dh_out = (sg > 0) ? carry[j][sg & 1] : nullptr;
                    rec_bwd(&a, 1, Dp, dp.dec_off, dp.dec_nact, q, t0, tl, j, want8[(size_t)j * nseg + sg], pad_wave, 0);   // the reverse wavefront pairs the same (j, sg) launches
                    done[j * nseg + sg] = next_event();
                    CUDA_CHECK(cudaEventRecord(done[j * nseg + sg], q));
                }
            }
            for (int j = 0; j < L; ++j) CUDA_CHECK(cudaStreamWaitEvent(s, done[j * nseg + 0], 0));
            kend();
        }
        if (side) side_after_main();   // the wavefront has joined the main stream: dGX/dGH/HP of all layers are final
        for (int j = L - 1; j >= 0; --j) {
            const std::string pre = "decode/rnn/l" + std::to_string(j) + "/";
            dec_wgrad(j, dGXl[j], dGHl[j], HPl[j], side ? swg : nullptr);
            RUN(launch_row_scatter(dh0l[j], H, dhx_sorted.f, H, nullptr, b, H, 1, s));
            const size_t end = pinfo(pre + "bR").off + align_up(3 * H, 64);
            allreduce_bucket(bucket_lo, end, side ? swg : nullptr);
            bucket_lo = end;
        }
        gemm(dGXl[0], 0, pmat("decode/rnn/l0/W"), 1, dYn, N, D, 3 * H, 1.f, nullptr, 0);
        std::swap(dY, dYn);
    }
    // d emb_tgt -> IndexedSlices part of dE (model.py:111)
    RUN(launch_embed_scatter_add(dp.ids_lead, N, dY.f, D, D, gE.f, s));
    phase("dec_bwd");

    // ---------------- backward: latent
    Mat dhx = both(b, D);
    RUN(launch_row_gather(dhx_sorted.f, nullptr, H, nullptr, dhx.f, dhx.h, D, dp.dec_perm, b, H, s));
    Mat dz = f32(b, R);
    gemm(dhx, 0, pmat("latent/ex/kernel"), 0, dz, b, R, D, 1.f, nullptr, 0);
    Mat dmulv = both(b, 2 * R);
    RUN(launch_latent_bwd(dz.f, mulv.f, eps_used, b, R, 1, anneal / ((float)b_glob * (float)R), dmulv.f, dmulv.h, s));
    Mat dhenc = f32(b, EH);
    Mat att_dK, att_dV;   // attentive: gradients of the key / value rows, folded into d hs below
    {   // the three affines' weight / bias gradients: behind the chain (side stream) when it is in use
        cudaStream_t qw = (side && !attentive) ? swg : nullptr;
        if (qw) side_after_main();
        gemm(z, 1, dhx, 1, gmat("latent/ex/kernel"), R, D, b, 1.f, nullptr, 1, qw);
        colsum(dhx, b, D, gptr("latent/ex/bias"), 0, qw);
        gemm(henc, 1, dmulv.colslice(0, R), 1, gmat("latent/mu/kernel"), EH, R, b, 1.f, nullptr, 1, qw);
        gemm(henc, 1, dmulv.colslice(R, R), 1, gmat("latent/lv/kernel"), EH, R, b, 1.f, nullptr, 1, qw);
        Mat a_mu(dmulv.f, nullptr, b, R, 2 * R), a_lv(dmulv.f + R, nullptr, b, R, 2 * R);
        colsum(a_mu, b, R, gptr("latent/mu/bias"), 0, qw);
        colsum(a_lv, b, R, gptr("latent/lv/bias"), 0, qw);
        gemm(dmulv.colslice(0, R), 0, pmat("latent/mu/kernel"), 0, dhenc, b, EH, R, 1.f, nullptr, 0);
        gemm(dmulv.colslice(R, R), 0, pmat("latent/lv/kernel"), 0, dhenc, b, EH, R, 1.f, nullptr, 1);
        size_t end = pinfo("latent/lv/bias").off + align_up(R, 64);
        if (attentive) {   // back through the layer norm, p, the softmax and q / k / v (all on the main stream)
            const Mat& HSx = encX[L];
            Mat dx = both(b, EH), dgr = f32(b, EH), dyy = f32(b, EH), dq = both(b, EH);
            att_dK = both(S, EH); att_dV = both(S, EH);
            RUN(launch_ln_bwd(dhenc.f, att.xhat, att.rstd, p + pinfo("encode/cata/LayerNorm/gamma").off, b, EH, dx.f, dx.h, dgr.f, s));
            colsum(dgr, b, EH, gptr("encode/cata/LayerNorm/gamma"), 0);
            colsum(dhenc, b, EH, gptr("encode/cata/LayerNorm/beta"), 0);
            gemm(att.y, 1, dx, 1, gmat(cata("p", "kernel")), EH, EH, b, 1.f, nullptr, 1);
            colsum(Mat(dx.f, nullptr, b, EH, EH), b, EH, gptr(cata("p", "bias")), 0);
            gemm(dx, 0, pmat(cata("p", "kernel")), 0, dyy, b, EH, EH, 1.f, nullptr, 0);
            RUN(launch_attn_bwd(dyy.f, att.q.f, att.K.f, att.Vv.f, att.prob, b, EH, ATT_HEADS, dp.enc_off, E.Tmax, dp.enc_last,
                                dq.f, dq.h, att_dK.f, att_dK.h, att_dV.f, att_dV.h, s));
            gemm(att.h0, 1, dq, 1, gmat(cata("q", "kernel")), EH, EH, b, 1.f, nullptr, 1);
            colsum(Mat(dq.f, nullptr, b, EH, EH), b, EH, gptr(cata("q", "bias")), 0);
            gemm(HSx, 1, att_dK, 1, gmat(cata("k", "kernel")), EH, EH, S, 1.f, nullptr, 1);
            colsum(Mat(att_dK.f, nullptr, S, EH, EH), S, EH, gptr(cata("k", "bias")), 0);
            gemm(HSx, 1, att_dV, 1, gmat(cata("v", "kernel")), EH, EH, S, 1.f, nullptr, 1);
            colsum(Mat(att_dV.f, nullptr, S, EH, EH), S, EH, gptr(cata("v", "bias")), 0);
            // d h (the gathered final state) = residual branch + query branch
            gemm(dq, 0, pmat(cata("q", "kernel")), 0, Mat(dx.f, nullptr, b, EH, EH), b, EH, EH, 1.f, nullptr, 1);
            dhenc = Mat(dx.f, nullptr, b, EH, EH);
            end = pinfo(cata("v", "bias")).off + align_up(EH, 64);
        }
        allreduce_bucket(bucket_lo, end, qw);
        bucket_lo = end;
    }
    phase("latent_bwd");

    // ---------------- backward: encoder (gather_nd adjoint, then BPTT through 3 x 2 GRUs)
    Mat dHS = f32(S, 2 * H), dHSn = f32(S, 2 * H);
    RUN(CUDA_CHECK(cudaMemsetAsync(dHS.f, 0, sizeof(float) * S * 2 * H, s)));
    if (enc_kind == 0) RUN(launch_row_scatter(dhenc.f, 2 * H, dHS.f, 2 * H, dp.enc_last, b, 2 * H, 0, s));
    auto attn_into = [&](const Mat& dTopRows) {   // every step's output also fed a key and a value
        if (!attentive) return;
        gemm(att_dK, 0, pmat(cata("k", "kernel")), 0, dTopRows, S, EH, EH, 1.f, nullptr, 1);
        gemm(att_dV, 0, pmat(cata("v", "kernel")), 0, dTopRows, S, EH, EH, 1.f, nullptr, 1);
    };
    if (enc_kind == 0) attn_into(dHS);
    size_t adam_split = 0;   // parameters [0, adam_split) were updated early on the side stream
    // Adam, TF-1 form (model.py:189): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), epsilon outside the bias-corrected root
    auto adam_range = [&](size_t lo, size_t hi, cudaStream_t q, const char* timer) {
        const double t = (double)(step + 1);
        const float lr_t = (float)((double)lr * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
        kbegin(timer, q);
        launch_adam(p + lo, g + lo, m + lo, v + lo, ph ? ph + lo : nullptr, (long long)(hi - lo), lr_t, 0.9f, 0.999f, 1e-8f, q);
        kend(q);
    };
    const bool adam_early = apply_update && L >= 2 && early_adam && !(dp_one_allreduce && cfg.nranks > 1);
    Mat dGXe1 = act(S, 6 * H), dGHe1 = act(S, 6 * H), HPe1 = act(S, 2 * H);
    // second set of gate-gradient buffers: layer i's weight-gradient GEMMs read one set on the side stream while layer
    // i-1's recurrence fills the other
    const bool side_enc = wgrad_overlap && use_mma && enc_kind == 0;
    Mat dGXe2 = side_enc ? act(S, 6 * H) : dGXe1, dGHe2 = side_enc ? act(S, 6 * H) : dGHe1, HPe2 = side_enc ? act(S, 2 * H) : HPe1;
    cudaEvent_t set_free[2] = {nullptr, nullptr};   // side stream done with the set (two layers ago)
    if (enc_kind != 0) {
        const Mat &dGXe = dGXe1, &dGHe = dGHe1, &HPe = HPe1;
        // BPTT through the independent stack(s): the top layer reads its slice of d hs, lower layers their own dX
        Mat dTop(dHS.f, nullptr, S, EH, EH);
        RUN(launch_row_scatter(dhenc.f, EH, dTop.f, EH, dp.enc_last, b, EH, 0, s));
        attn_into(dTop);
        Mat dcur[2], dnext[2];
        for (int d = 0; d < nd_enc; ++d) { dcur[d] = f32(S, H); dnext[d] = f32(S, H); }
        Mat dX0(dHSn.f, nullptr, S, D, D);
        for (int j = L - 1; j >= 0; --j) {
            GruBwdArgs a[2];
            for (int d = 0; d < nd_enc; ++d) {
                const std::string pre = enc_prefix(d, j);
                const Mat& out = encIn[d][j + 1];
                if (j == L - 1) { a[d].dhs = dTop.f + d * H; a[d].ld_dhs = EH; }
                else { a[d].dhs = dcur[d].f; a[d].ld_dhs = H; }
                a[d].hs_f = out.f; a[d].hs_h = out.h; a[d].ld_hs = out.ld;
                a[d].h0 = nullptr;
                a[d].cache = encCache[d * L + j];
                a[d].R_f = p + pinfo(pre + "R").off;
                a[d].R_h = ph ? ph + pinfo(pre + "R").off : nullptr;
                a[d].dgx_f = dGXe.f ? dGXe.f + d * 3 * H : nullptr;
                a[d].dgx_h = dGXe.h ? dGXe.h + d * 3 * H : nullptr;
                a[d].dgh_f = dGHe.f ? dGHe.f + d * 3 * H : nullptr;
                a[d].dgh_h = dGHe.h ? dGHe.h + d * 3 * H : nullptr;
                a[d].ld_dg = 6 * H;
                a[d].hp_f = HPe.f ? HPe.f + d * H : nullptr;
                a[d].hp_h = HPe.h ? HPe.h + d * H : nullptr;
                a[d].ld_hp = 2 * H;
                a[d].dh0 = nullptr;
                a[d].reverse = (enc_kind == 1 && d == 1) ? 1 : 0;
            }
            gru_bwd(a, nd_enc, E, dp.enc_off, dp.enc_nact);
            for (int d = 0; d < nd_enc; ++d) {
                const std::string pre = enc_prefix(d, j);
                const int in = (j == 0) ? D : H;
                Mat dgx = dGXe.colslice(d * 3 * H, 3 * H), dgh = dGHe.colslice(d * 3 * H, 3 * H);
                gemm(dgx, 1, encIn[d][j], 1, gmat(pre + "W"), 3 * H, in, S, 1.f, nullptr, 1);
                gemm(dgh, 1, HPe.colslice(d * H, H), 1, gmat(pre + "R"), 3 * H, H, S, 1.f, nullptr, 1);
                colsum(dgx, S, 3 * H, gptr(pre + "bW"));
                colsum(dgh, S, 3 * H, gptr(pre + "bR"));
                if (j > 0) gemm(dgx, 0, pmat(pre + "W"), 1, dnext[d], S, H, 3 * H, 1.f, nullptr, 0);
                else gemm(dgx, 0, pmat(pre + "W"), 1, dX0, S, D, 3 * H, 1.f, nullptr, d > 0 ? 1 : 0);   // both stacks read emb_src
                std::swap(dcur[d], dnext[d]);
            }
            const size_t end = pinfo(enc_prefix(nd_enc - 1, j) + "bR").off + align_up(3 * H, 64);
            allreduce_bucket(bucket_lo, end);
            bucket_lo = end;
        }
        std::swap(dHS, dHSn);   // dHS now holds d emb_src (S,D) like the stacked branch leaves it
    }
    for (int i = L - 1; i >= 0 && enc_kind == 0; --i) {
        const int in = (i == 0) ? D : 2 * H;
        const std::string pre = "encode/rnn" + std::to_string(i + 1) + "/";
        const int set = side_enc ? ((L - 1 - i) & 1) : 0;
        const Mat& dGXe = set ? dGXe2 : dGXe1;
        const Mat& dGHe = set ? dGHe2 : dGHe1;
        const Mat& HPe = set ? HPe2 : HPe1;
        if (side && set_free[set]) CUDA_CHECK(cudaStreamWaitEvent(s, set_free[set], 0));
        // last layer: nothing below it to hide its weight gradients behind, so they are taken per time segment
        const bool seg_wgrad = side && side_enc && i == 0 && enc_segmented(E) && seg_wgrad_on;
        GruBwdArgs a[2];
        for (int d = 0; d < 2; ++d) {
            const std::string pd = pre + (d ? "bwd/" : "fwd/");
            a[d].dhs = dHS.f + d * H; a[d].ld_dhs = 2 * H;
            a[d].hs_f = encX[i + 1].f ? encX[i + 1].f + d * H : nullptr;
            a[d].hs_h = encX[i + 1].h ? encX[i + 1].h + d * H : nullptr;
            a[d].ld_hs = 2 * H;
            a[d].h0 = nullptr;
            a[d].cache = encCache[2 * i + d];
            a[d].R_f = p + pinfo(pd + "R").off;
            a[d].R_h = ph ? ph + pinfo(pd + "R").off : nullptr;
            a[d].dgx_f = dGXe.f ? dGXe.f + d * 3 * H : nullptr;
            a[d].dgx_h = dGXe.h ? dGXe.h + d * 3 * H : nullptr;
            a[d].dgh_f = dGHe.f ? dGHe.f + d * 3 * H : nullptr;
            a[d].dgh_h = dGHe.h ? dGHe.h + d * 3 * H : nullptr;
            a[d].ld_dg = 6 * H;
            a[d].hp_f = HPe.f ? HPe.f + d * H : nullptr;
            a[d].hp_h = HPe.h ? HPe.h + d * H : nullptr;
            a[d].ld_hp = 2 * H;
            a[d].dh0 = nullptr;
            a[d].reverse = d;
        }
        if (enc_segmented(E)) {
            // BPTT as two chains of segment launches: the forward GRU's walks the segments downwards, the reverse GRU's
            // upwards; the gradient wrt the carried state is handed from launch to launch
            const int nsegE = (E.Tmax + enc_seg - 1) / enc_seg;
            float* carry[2][2];
            for (int d = 0; d < 2; ++d)
                for (int q = 0; q < 2; ++q) carry[d][q] = (float*)arena.alloc(sizeof(float) * b * H);
            if (!arena.dry) {
                std::vector<int> want8;
                enc_slice_plan(E, nsegE, true, &want8);
                for (int d = 0; d < 2; ++d)
                    for (int q = 0; q < 2; ++q) CUDA_CHECK(cudaMemsetAsync(carry[d][q], 0, sizeof(float) * b * H, s));
                kbegin("k:gru_bwd_enc");
                cudaEvent_t fork = next_event();
                CUDA_CHECK(cudaEventRecord(fork, s));
                for (int d = 0; d < 2; ++d) CUDA_CHECK(cudaStreamWaitEvent(sw[d], fork, 0));
                for (int k = 0; k < nsegE; ++k)
                    for (int d = 0; d < 2; ++d) {
                        const int sg = d ? k : nsegE - 1 - k;
                        const int t0 = sg * enc_seg, tl = std::min(enc_seg, E.Tmax - t0);
                        GruBwdArgs x = a[d];
                        x.dh_in = (k > 0) ? carry[d][(k - 1) & 1] : nullptr;
                        x.dh_out = (k + 1 < nsegE) ? carry[d][k & 1] : nullptr;
                        rec_bwd(&x, 1, E, dp.enc_off, dp.enc_nact, sw[d], t0, tl, d, want8[(size_t)d * nsegE + k], pad_wave, enc_bwd_chunk);
                        if (seg_wgrad) {
                            // the segment's rows of d gates are final: their share of the weight / bias gradients goes to
                            // the side stream now, so only the last segment's share is left when the chains end
                            cudaEvent_t ev = next_event();
                            CUDA_CHECK(cudaEventRecord(ev, sw[d]));
                            CUDA_CHECK(cudaStreamWaitEvent(swg, ev, 0));
                            const std::string pd = pre + (d ? "bwd/" : "fwd/");
                            const long long r0 = E.off[t0], nr = E.off[t0 + tl] - r0;
                            Mat dgx = dGXe.colslice(d * 3 * H, 3 * H).rowslice(r0, nr), dgh = dGHe.colslice(d * 3 * H, 3 * H).rowslice(r0, nr);
                            gemm(dgx, 1, encX[i].rowslice(r0, nr), 1, gmat(pd + "W"), 3 * H, in, nr, 1.f, nullptr, 1, swg);
                            gemm(dgh, 1, HPe.colslice(d * H, H).rowslice(r0, nr), 1, gmat(pd + "R"), 3 * H, H, nr, 1.f, nullptr, 1, swg);
                            colsum(dgx, nr, 3 * H, gptr(pd + "bW"), 1, swg);
                            colsum(dgh, nr, 3 * H, gptr(pd + "bR"), 1, swg);
                        }
                    }
                for (int d = 0; d < 2; ++d) {
                    cudaEvent_t ev = next_event();
                    CUDA_CHECK(cudaEventRecord(ev, sw[d]));
                    CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
                }
                kend();
            }
        } else {
            gru_bwd(a, 2, E, dp.enc_off, dp.enc_nact);
        }
        // the layer below waits for d x only; with the side stream the weight gradients follow behind it. The last
        // layer's have nothing left to hide behind and stay on the main stream.
        const bool on_side = side && side_enc && i > 0;
        cudaStream_t qw = on_side ? swg : nullptr;
        Mat W = pmat(pre + "fwd/W");
        W.rows = 6 * H;
        Mat dX(dHSn.f, nullptr, S, in, in);
        if (on_side) {
            side_after_main();   // the recurrence has joined the main stream
            gemm(dGXe, 0, W, 1, dX, S, in, 6 * H, 1.f, nullptr, 0);
        }
        Mat gW = gmat(pre + "fwd/W");
        gW.rows = 6 * H;
        if (!seg_wgrad) {
            gemm(dGXe, 1, encX[i], 1, gW, 6 * H, in, S, 1.f, nullptr, 1, qw);
            for (int d = 0; d < 2; ++d) {
                const std::string pd = pre + (d ? "bwd/" : "fwd/");
                gemm(dGHe.colslice(d * 3 * H, 3 * H), 1, HPe.colslice(d * H, H), 1, gmat(pd + "R"), 3 * H, H, S, 1.f, nullptr, 1, qw);
            }
            colsum(dGXe, S, 6 * H, gptr(pre + "fwd/bW"), 0, qw);   // fwd/bW and bwd/bW are adjacent
            colsum(dGHe, S, 6 * H, gptr(pre + "fwd/bR"), 0, qw);
        } else {
            qw = swg;   // the layer's all-reduce bucket follows the side stream
        }
        if (!on_side) gemm(dGXe, 0, W, 1, dX, S, in, 6 * H, 1.f, nullptr, 0);
        std::swap(dHS, dHSn);
        if (on_side) {
            set_free[set] = next_event();
            CUDA_CHECK(cudaEventRecord(set_free[set], swg));
        }
        const size_t end = pinfo(pre + "bwd/bR").off + align_up(3 * H, 64);
        allreduce_bucket(bucket_lo, end, qw);
        bucket_lo = end;
        if (on_side && i == 1 && adam_early) {
            // Every parameter in front of the first encoder layer's has its final gradient once the side stream (data
            // parallel: the NCCL stream, behind the layer-2 bucket's all-reduce) gets here, and no reader left on the
            // chain after the dgrad above: their Adam update (70 % of the 683 MB the update moves) runs under the last
            // layer's recurrence instead of after it.
            // single GPU: on the side stream (behind the layer-2 weight gradients).  Data parallel: on a stream of its own
            // that waits for the layer-2 bucket's all-reduce -- NOT on the NCCL stream, where the 0.3 ms this bandwidth-bound
            // kernel takes on the few SMs the recurrence leaves free would sit in front of the last buckets' all-reduces
            cudaStream_t qa = cfg.nranks > 1 ? sad : swg;
            cudaEvent_t ev = next_event();
            CUDA_CHECK(cudaEventRecord(ev, s));
            CUDA_CHECK(cudaStreamWaitEvent(qa, ev, 0));
            if (cfg.nranks > 1) {
                cudaEvent_t evc = next_event();
                CUDA_CHECK(cudaEventRecord(evc, st[2]));
                CUDA_CHECK(cudaStreamWaitEvent(qa, evc, 0));
            }
            adam_split = end;
            adam_range(0, adam_split, qa, "k:adam_side");
        }
    }
    if (side) {   // everything the side stream produced is a gradient: Adam and the last bucket come after it
        cudaEvent_t ev = next_event();
        CUDA_CHECK(cudaEventRecord(ev, swg));
        CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
    }
    // d emb_src -> IndexedSlices part of dE (model.py:112); dHS now holds (S,D) with ld D
    RUN(launch_embed_scatter_add(dp.ids_src, S, dHS.f, D, D, gE.f, s));
    phase("enc_bwd");
    allreduce_bucket(bucket_lo, nflat);
    bucket_lo = nflat;
    if (cfg.nranks > 1 && !arena.dry) {
        CUDA_CHECK(cudaEventRecord(ev_comm, st[2]));
        CUDA_CHECK(cudaStreamWaitEvent(s, ev_comm, 0));
        if (adam_split) {   // the early Adam ran on its own stream behind the layer-2 bucket's all-reduce
            cudaEvent_t ev = next_event();
            CUDA_CHECK(cudaEventRecord(ev, sad));
            CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
        }
    }
    phase("allreduce_wait");

    // ---------------- Adam, TF-1 form (model.py:189)
    if (apply_update) RUN(adam_range(adam_split, nflat, s, "k:adam"));
    phase("adam");
}

// ------------------------------------------------------------------------------------------
// public steps
// ------------------------------------------------------------------------------------------
void Engine::train_step_submit(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, const uint8_t* keep, const float* eps,
                               int64_t n_tok_global, int64_t b_global, int64_t row0, bool apply_update) {
    if (pend_n >= 2) throw std::runtime_error("train_step_submit: two steps are already in flight, call train_step_wait first");
    PendingStep& ps = pend[submit_seq & 1];
    if (!ps.done) {
        CUDA_CHECK(cudaEventCreateWithFlags(&ps.done, cudaEventDisableTiming));
        CUDA_CHECK(cudaMallocHost(&ps.h_stats, 4 * sizeof(double)));
    }
    schedule_f32(step, cfg.accelerate, cfg.learn_rate, &ps.keepwd, &ps.anneal, &ps.lr);
    DropoutSpec drop;
    drop.train = 1; drop.keep = keep; drop.rate_keepwd = ps.keepwd; drop.seed = seed; drop.step = (uint64_t)step; drop.row0 = row0;
    last.train = 1; last.n_tok_global = n_tok_global; last.b_global = b_global; last.row0 = row0;
    if (!next_rows.empty()) {
        if ((int)next_rows.size() != b) {
            next_rows.clear();
            throw std::runtime_error("train step: argsim_set_global_rows was given another number of rows than this batch has");
        }
        drop.rows = next_rows.data();
    }
    stage(src, tgt, b, Ts, Tt, 1, drop, eps);
    next_rows.clear();
    run_device(2, apply_update);
    CUDA_CHECK(cudaMemcpyAsync(ps.h_stats, d_stats, 4 * sizeof(double), cudaMemcpyDeviceToHost, st[0]));
    CUDA_CHECK(cudaEventRecord(ps.done, st[0]));   // the comm and side streams have joined the main stream by now
    ps.n_glob = n_tok_global > 0 ? (double)n_tok_global : (double)plan.dec.rows;
    ps.b_glob = b_global > 0 ? (double)b_global : (double)b;
    if (apply_update) step += 1;
    ps.step_after = step;
    ++submit_seq;
    ++pend_n;
}

void Engine::train_step_wait(argsim_step_stats* out) {
    if (pend_n == 0) throw std::runtime_error("train_step_wait: no step in flight");
    PendingStep& ps = pend[(submit_seq - pend_n) & 1];
    CUDA_CHECK(cudaEventSynchronize(ps.done));
    --pend_n;
    if (pend_n == 0) {   // the timers' events belong to the newest submission
        CUDA_CHECK(cudaStreamSynchronize(st[2]));
        collect_timings();
    }
    if (out) {
        out->loss_gen = (float)(ps.h_stats[0] / ps.n_glob);
        out->errt = (float)(ps.h_stats[1] / ps.n_glob);
        out->loss_kld = (float)(ps.h_stats[2] / (ps.b_glob * R));
        out->loss = ps.anneal * out->loss_kld + out->loss_gen;
        out->rate_keepwd = ps.keepwd; out->rate_anneal = ps.anneal; out->rate_update = ps.lr;
        out->n_tokens = (int64_t)ps.n_glob;
        out->step = ps.step_after;
    }
}

void Engine::set_global_rows(const int64_t* rows, int b) {
    next_rows.clear();
    if (rows && b > 0) next_rows.assign(rows, rows + b);
    for (int64_t r : next_rows)
        if (r < 0 || r > 0x7fffffffLL) {
            next_rows.clear();
            throw std::runtime_error("global row indices must be in [0, 2^31)");
        }
}

void Engine::drain() {
    for (int i = pend_n; i > 0; --i) CUDA_CHECK(cudaEventSynchronize(pend[(submit_seq - i) & 1].done));
}

void Engine::train_step(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, const uint8_t* keep, const float* eps,
                        int64_t n_tok_global, int64_t b_global, int64_t row0, bool apply_update, argsim_step_stats* out) {
    if (pend_n) throw std::runtime_error("train_step: pipelined steps are in flight, call train_step_wait first");
    train_step_submit(src, tgt, b, Ts, Tt, keep, eps, n_tok_global, b_global, row0, apply_update);
    train_step_wait(out);
}

void Engine::bench_resident(int iters, float* ms) {
    if (!last.train) throw std::runtime_error("bench_resident: no staged training batch (call argsim_train_step first)");
    cudaEvent_t a, b2;
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b2));
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    CUDA_CHECK(cudaEventRecord(a, st[0]));
    for (int i = 0; i < iters; ++i) {
        run_device(2, true);
        step += 1;
    }
    CUDA_CHECK(cudaEventRecord(b2, st[0]));
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    CUDA_CHECK(cudaStreamSynchronize(st[2]));
    float t = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&t, a, b2));
    collect_timings();
    *ms = t / (float)std::max(iters, 1);
    cudaEventDestroy(a); cudaEventDestroy(b2);
}

void Engine::eval_step(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, float* errt, float* lgen, int64_t cap,
                       float* lkld, int64_t* n_rows, int32_t* pred) {
    DropoutSpec drop;
    last = StepArgs();
    stage(src, tgt, b, Ts, Tt, 1, drop, nullptr);
    const long long N = plan.dec.rows;
    if (n_rows) *n_rows = N;
    if ((errt || lgen || pred) && cap < N) throw std::runtime_error("eval_step: per-row output capacity too small");
    run_device(1, false);
    size_t need = (size_t)N * 3 + (size_t)b * R;
    if (need > h_out_cap) {
        cudaFreeHost(h_out);
        CUDA_CHECK(cudaMallocHost(&h_out, need * sizeof(float)));
        h_out_cap = need;
    }
    float* hl = h_out; float* he = h_out + N; int* hp = (int*)(h_out + 2 * N); float* hk = h_out + 3 * N;
    CUDA_CHECK(cudaMemcpyAsync(hl, outp.loss_samp, N * sizeof(float), cudaMemcpyDeviceToHost, st[0]));
    CUDA_CHECK(cudaMemcpyAsync(he, outp.err_samp, N * sizeof(float), cudaMemcpyDeviceToHost, st[0]));
    CUDA_CHECK(cudaMemcpyAsync(hp, outp.pred, N * sizeof(int), cudaMemcpyDeviceToHost, st[0]));
    CUDA_CHECK(cudaMemcpyAsync(hk, outp.kld_samp, (size_t)b * R * sizeof(float), cudaMemcpyDeviceToHost, st[0]));
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    collect_timings();
    // back to the reference's boolean_mask order
    for (long long r = 0; r < N; ++r) {
        const long long q = plan.ref_row[r];
        if (lgen) lgen[q] = hl[r];
        if (errt) errt[q] = he[r];
        if (pred) pred[q] = hp[r];
    }
    if (lkld) memcpy(lkld, hk, (size_t)b * R * sizeof(float));
}

void Engine::embed_one(const int32_t* src, int b, int T, float* mu_out) {
    DropoutSpec drop;
    last = StepArgs();
    stage(src, nullptr, b, T, 0, 0, drop, nullptr);
    run_device(0, false);
    // mulv is (b, 2R) with only the mu half written
    CUDA_CHECK(cudaMemcpy2DAsync(mu_out, (size_t)R * sizeof(float), outp.mulv, (size_t)2 * R * sizeof(float),
                                 (size_t)R * sizeof(float), b, cudaMemcpyDeviceToHost, st[0]));
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    collect_timings();
}

// Rows are independent in 'infer' mode, so a batch larger than the persistent recurrence holds resident
// (4 slices x 128 rows) is processed as length-sorted micro-batches: each one's step count is its own longest
// row, not the batch maximum (eval_embed*.py feed 128 rows per call; BASELINE configs[3] feeds 4096).
void Engine::embed(const int32_t* src, int b, int T, float* mu_out) {
    // micro-batch: 512 rows fill the mma.sync kernel's 4 slices x 128 rows per direction; the tensor-memory kernel holds 256
    // rows per slice (4 chunks of 64 in flight through its two operand slots)
    static const int cap_env = getenv("ARGSIM_EMBED_CAP") ? atoi(getenv("ARGSIM_EMBED_CAP")) : 0;
    const int cap = cap_env > 0 ? cap_env : (tc && (gru_tc_mode & 5)) ? 1024 : 512;
    if (!use_mma || b <= cap) {
        embed_one(src, b, T, mu_out);
        return;
    }
    std::vector<int> len(b), order(b);
    for (int i = 0; i < b; ++i) {
        int n = 0;
        for (int t = 0; t < T; ++t) n += (src[(size_t)i * T + t] != cfg.eos);
        len[i] = n;
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return len[x] > len[y]; });
    // The micro-batches are pipelined: the host plan + H2D of micro-batch i+1 are issued while the device runs i, every
    // mu lands in one pinned buffer and the stream is synchronised once at the end (two staging slots, one event each)
    const size_t need = (size_t)b * R;
    if (need > h_out_cap) {
        CUDA_CHECK(cudaStreamSynchronize(st[0]));
        cudaFreeHost(h_out);
        CUDA_CHECK(cudaMallocHost(&h_out, need * sizeof(float)));
        h_out_cap = need;
    }
    cudaEvent_t ev_slot[2] = {next_event_persistent(0), next_event_persistent(1)};
    std::vector<int32_t> rows;
    DropoutSpec drop;
    last = StepArgs();
    int mb = 0;
    try {
        for (int i0 = 0; i0 < b; i0 += cap, ++mb) {
            const int n = std::min(cap, b - i0);
            const int Tm = std::max(1, len[order[i0]]);     // longest row of this micro-batch (sorted descending)
            rows.assign((size_t)n * Tm, cfg.eos);
            for (int r = 0; r < n; ++r) memcpy(rows.data() + (size_t)r * Tm, src + (size_t)order[i0 + r] * T, sizeof(int32_t) * std::min(Tm, T));
            // a row with eos inside its first Tm tokens keeps it and fails the trim() contract check, as it must
            stage_slot = mb & 1;
            if (mb >= 2) CUDA_CHECK(cudaEventSynchronize(ev_slot[stage_slot]));   // the slot's previous H2D copy has been consumed
            stage(rows.data(), nullptr, n, Tm, 0, 0, drop, nullptr);
            run_device(0, false);
            CUDA_CHECK(cudaMemcpy2DAsync(h_out + (size_t)i0 * R, (size_t)R * sizeof(float), outp.mulv, (size_t)2 * R * sizeof(float),
                                         (size_t)R * sizeof(float), n, cudaMemcpyDeviceToHost, st[0]));
            CUDA_CHECK(cudaEventRecord(ev_slot[stage_slot], st[0]));
            if (mb >= 1) {   // the previous micro-batch's mu has landed: hand its rows out while the device runs this one
                CUDA_CHECK(cudaEventSynchronize(ev_slot[(mb - 1) & 1]));
                for (int r = i0 - cap; r < i0; ++r) memcpy(mu_out + (size_t)order[r] * R, h_out + (size_t)r * R, sizeof(float) * R);
            }
        }
    } catch (...) {
        stage_slot = -1;
        cudaStreamSynchronize(st[0]);
        throw;
    }
    stage_slot = -1;
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    collect_timings();
    for (int r = (mb - 1) * cap; r < b; ++r) memcpy(mu_out + (size_t)order[r] * R, h_out + (size_t)r * R, sizeof(float) * R);
}

// decode(), model.py:204-219.  FP32_VALIDATE handles run fp32 SIMT GEMMs (bit-exact tokens against the oracle); BF16
// handles run the step on the tensor cores: per layer one tcgen05 GEMM for W x + bW and the fused per-step recurrence
// kernel (R streamed by TMA, tcgen05 SS form, gate epilogue: k_gru_step_fwd), then the out projection and the tied
// vocabulary GEMM on k_gemm_tc2 and the arg-max kernel -- 11 launches per token, no host round trip in decode_loop.
struct Engine::DecodeBufs {
    int b = 0, out_steps = 0;
    bool tc = false;
    float *dz = nullptr, *st_f = nullptr, *x = nullptr, *gx = nullptr, *gh = nullptr, *ho = nullptr, *lg = nullptr;
    ::bf16 *st_h = nullptr, *x_h = nullptr, *ho_h = nullptr, *z_h = nullptr;   // bf16 path: state ping-pong (2,L,b,H)
    int *lead = nullptr, *pred = nullptr, *out = nullptr, *flag = nullptr;
    int cur = 0;
    cudaGraphExec_t two_steps = nullptr;   // bf16 path: two tokens (one ping-pong period) as one replayable graph
    void release() {
        if (two_steps) cudaGraphExecDestroy(two_steps);
        two_steps = nullptr;
        cudaFree(dz); cudaFree(st_f); cudaFree(x); cudaFree(gx); cudaFree(gh); cudaFree(ho); cudaFree(lg);
        cudaFree(st_h); cudaFree(x_h); cudaFree(ho_h); cudaFree(z_h); cudaFree(lead); cudaFree(pred); cudaFree(out); cudaFree(flag);
        dz = st_f = x = gx = gh = ho = lg = nullptr; st_h = x_h = ho_h = z_h = nullptr; lead = pred = out = flag = nullptr;
        b = out_steps = 0;
    }
    ~DecodeBufs() { release(); }
};

// decode buffers live with the engine: reallocated only when the batch (or the step budget of the loop) changes
Engine::DecodeBufs& Engine::decode_bufs(int b, int steps) {
    if (!dbuf) dbuf = new DecodeBufs();
    DecodeBufs& B = *dbuf;
    if (B.b == b && B.out_steps >= steps) return B;
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    if (B.b == b) {   // a longer step budget: only the token buffer grows (the graph holds its address)
        if (B.two_steps) cudaGraphExecDestroy(B.two_steps);
        B.two_steps = nullptr;
        cudaFree(B.out);
        B.out = nullptr;
        CUDA_CHECK(cudaMalloc(&B.out, sizeof(int) * (size_t)b * steps));
        B.out_steps = steps;
        return B;
    }
    const int keep_steps = steps;
    B.release();
    B.b = b;
    B.out_steps = keep_steps;
    B.tc = use_tc && gru_step_supported(H, b);
    CUDA_CHECK(cudaMalloc(&B.dz, sizeof(float) * b * R));
    CUDA_CHECK(cudaMalloc(&B.st_f, sizeof(float) * (L + 1) * b * H));
    CUDA_CHECK(cudaMalloc(&B.gx, sizeof(float) * b * 3 * H));
    CUDA_CHECK(cudaMalloc(&B.ho, sizeof(float) * b * D));
    CUDA_CHECK(cudaMalloc(&B.lg, sizeof(float) * b * V));
    CUDA_CHECK(cudaMalloc(&B.lead, sizeof(int) * b));
    CUDA_CHECK(cudaMalloc(&B.pred, sizeof(int) * b));
    CUDA_CHECK(cudaMalloc(&B.flag, sizeof(int) * 4));
    if (keep_steps > 0) CUDA_CHECK(cudaMalloc(&B.out, sizeof(int) * (size_t)b * keep_steps));
    if (B.tc) {
        CUDA_CHECK(cudaMalloc(&B.st_h, sizeof(::bf16) * 2 * L * b * H));
        CUDA_CHECK(cudaMalloc(&B.x_h, sizeof(::bf16) * b * D));
        CUDA_CHECK(cudaMalloc(&B.ho_h, sizeof(::bf16) * b * D));
        CUDA_CHECK(cudaMalloc(&B.z_h, sizeof(::bf16) * b * R));
    } else {
        CUDA_CHECK(cudaMalloc(&B.x, sizeof(float) * b * D));
        CUDA_CHECK(cudaMalloc(&B.gh, sizeof(float) * b * 3 * H));
    }
    return B;
}

// state_in = stack((ex(z),) * L)  (model.py:156-158); dz holds z
void Engine::decode_seed(DecodeBufs& B) {
    cudaStream_t s = st[0];
    const int b = B.b;
    struct Guard { bool& f; Guard(bool& x) : f(x) { f = true; } ~Guard() { f = false; } } guard(decoding);
    float* dh = B.st_f + (size_t)L * b * H;
    if (B.tc) {
        launch_cast_bf16(B.dz, B.z_h, (long long)b * R, s);
        gemm(Mat(B.dz, B.z_h, b, R, R), 0, pmat("latent/ex/kernel"), 1, Mat(dh, nullptr, b, H, H), b, D, R, 1.f,
             p + pinfo("latent/ex/bias").off, 0);
    } else {
        gemm_simt(B.dz, R, 0, p + pinfo("latent/ex/kernel").off, D, 1, dh, H, b, D, R, 1.f, p + pinfo("latent/ex/bias").off, 0, nullptr, s);
    }
    for (int l = 0; l < L; ++l)
        CUDA_CHECK(cudaMemcpyAsync(B.st_f + (size_t)l * b * H, dh, sizeof(float) * b * H, cudaMemcpyDeviceToDevice, s));
    decode_state_changed(B);
}
// the fp32 state (L,b,H) was written from outside the step: refresh the bf16 copy the tensor-core step reads
void Engine::decode_state_changed(DecodeBufs& B) {
    if (!B.tc) return;
    B.cur = 0;
    launch_cast_bf16(B.st_f, B.st_h, (long long)L * B.b * H, st[0]);
}

// one token: lead (b) -> state (L,b,H) updated in place, pred (b) = arg-max of the logits (lowest index on ties)
void Engine::decode_one(DecodeBufs& B) {
    cudaStream_t s = st[0];
    const int b = B.b;
    struct Guard { bool& f; Guard(bool& x) : f(x) { f = true; } ~Guard() { f = false; } } guard(decoding);
    if (B.tc) {
        launch_embed_gather_bf16(B.lead, b, ph + pinfo("embed/embedding").off, D, B.x_h, s);
        const ::bf16* in = B.x_h;
        ::bf16* cur = B.st_h + (size_t)B.cur * L * b * H;
        ::bf16* nxt = B.st_h + (size_t)(1 - B.cur) * L * b * H;
        for (int l = 0; l < L; ++l) {
            const std::string pre = "decode/rnn/l" + std::to_string(l) + "/";
            gemm(Mat(nullptr, const_cast<::bf16*>(in), b, D, D), 0, pmat(pre + "W"), 0, Mat(B.gx, nullptr, b, 3 * H, 3 * H), b, 3 * H, D, 1.f,
                 p + pinfo(pre + "bW").off, 0);
            GruFwdArgs a{};
            a.gx = B.gx; a.ld_gx = 3 * H;
            a.R_f = p + pinfo(pre + "R").off; a.R_h = ph + pinfo(pre + "R").off; a.bR = p + pinfo(pre + "bR").off;
            a.ld_hs = H;
            float* stf = B.st_f + (size_t)l * b * H;
            ::bf16* hc = cur + (size_t)l * b * H;
            ::bf16* hn = nxt + (size_t)l * b * H;
            const long long zero = 0;
            gru_step_fwd(&a, 1, &b, &zero, b, H, &stf, &hc, &hn, s, true);
            in = hn;
        }
        B.cur ^= 1;
        gemm(Mat(nullptr, const_cast<::bf16*>(in), b, H, H), 0, pmat("decode/out/kernel"), 1, Mat(B.ho, B.ho_h, b, D, D), b, D, D, 1.f,
             p + pinfo("decode/out/bias").off, 0);
        const Mat HO(B.ho, B.ho_h, b, D, D), LG(B.lg, nullptr, b, V, V);
        if (tied) gemm(HO, 0, pmat("embed/embedding"), 0, LG, b, V, D, 1.0f / sqrtf((float)D), nullptr, 0);
        else gemm(HO, 0, pmat("logits/dense/kernel"), 1, LG, b, V, D, 1.f, p + pinfo("logits/dense/bias").off, 0);
    } else {
        launch_embed_gather_f32(B.lead, b, p + pinfo("embed/embedding").off, D, B.x, s);
        const float* in = B.x;
        for (int l = 0; l < L; ++l) {
            const std::string pre = "decode/rnn/l" + std::to_string(l) + "/";
            gemm_simt(in, D, 0, p + pinfo(pre + "W").off, D, 0, B.gx, 3 * H, b, 3 * H, D, 1.f, p + pinfo(pre + "bW").off, 0, nullptr, s);
            float* stl = B.st_f + (size_t)l * b * H;
            gru_generic_cell(B.gx, 3 * H, p + pinfo(pre + "R").off, p + pinfo(pre + "bR").off, stl, B.gh, b, H, s);
            in = stl;
        }
        gemm_simt(in, H, 0, p + pinfo("decode/out/kernel").off, D, 1, B.ho, D, b, D, D, 1.f, p + pinfo("decode/out/bias").off, 0, nullptr, s);
        if (tied) gemm_simt(B.ho, D, 0, p + pinfo("embed/embedding").off, D, 0, B.lg, V, b, V, D, 1.0f / sqrtf((float)D), nullptr, 0, nullptr, s);
        else gemm_simt(B.ho, D, 0, p + pinfo("logits/dense/kernel").off, V, 1, B.lg, V, b, V, D, 1.f, p + pinfo("logits/dense/bias").off, 0, nullptr, s);
    }
    launch_ce_f32(B.lg, V, nullptr, b, V, 0.f, 0, nullptr, nullptr, B.pred, d_stats, s);
}

void Engine::decode_init(const float* z, int b, float* state) {
    if (b <= 0) throw std::runtime_error("decode_init: b must be positive");
    DecodeBufs& B = decode_bufs(b, 0);
    copy_sync(B.dz, z, sizeof(float) * b * R, cudaMemcpyHostToDevice);
    decode_seed(B);
    CUDA_CHECK(cudaMemcpyAsync(state, B.st_f, sizeof(float) * L * b * H, cudaMemcpyDeviceToHost, st[0]));
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
}

void Engine::decode_step(const int32_t* lead, int b, float* state, int32_t* pred) {
    if (b <= 0) throw std::runtime_error("decode_step: b must be positive");
    for (int i = 0; i < b; ++i)
        if (lead[i] < 0 || lead[i] >= V) throw std::runtime_error("decode_step: token id out of range");
    cudaStream_t s = st[0];
    DecodeBufs& B = decode_bufs(b, 0);
    CUDA_CHECK(cudaMemcpyAsync(B.lead, lead, sizeof(int) * b, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(B.st_f, state, sizeof(float) * L * b * H, cudaMemcpyHostToDevice, s));
    decode_state_changed(B);
    decode_one(B);
    CUDA_CHECK(cudaMemcpyAsync(pred, B.pred, sizeof(int) * b, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(state, B.st_f, sizeof(float) * L * b * H, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
}

// decode() of the reference as ONE device-resident loop: no host round trip per token (the
// reference does a sess.run per step, model.py:215); the step counter, the all-eos stop test and the step budget live on
// the device and the host looks at the stop flag every 16 steps.  The same step as decode_step, so both paths give
// identical tokens.  BF16 handles replay two steps (one period of the state ping-pong) as one CUDA graph -- 24 kernel
// nodes per replay instead of 24 host launches (ARGSIM_DECODE_NO_GRAPH=1: plain launches).
int Engine::decode_loop(const float* z, int b, int steps, int32_t* tokens) {
    if (b <= 0 || steps <= 0) throw std::runtime_error("decode: b and steps must be positive");
    cudaStream_t s = st[0];
    DecodeBufs& B = decode_bufs(b, steps);
    const int init_flag[4] = {0, 0, steps, 0};
    CUDA_CHECK(cudaMemcpyAsync(B.flag, init_flag, sizeof(init_flag), cudaMemcpyHostToDevice, s));
    copy_sync(B.dz, z, sizeof(float) * b * R, cudaMemcpyHostToDevice);
    decode_seed(B);
    launch_fill_i32(B.lead, b, cfg.bos, s);
    int flag[4] = {0, 0, 0, 0};
    static const bool prof = getenv("ARGSIM_DECODE_PROF") != nullptr;   // device time of the loop against the host's issue time
    static const bool no_graph = getenv("ARGSIM_DECODE_NO_GRAPH") != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::chrono::steady_clock::time_point h0;
    if (prof) {
        CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
        CUDA_CHECK(cudaStreamSynchronize(s));
        CUDA_CHECK(cudaEventRecord(e0, s));
        h0 = std::chrono::steady_clock::now();
    }
    const bool graph = B.tc && !no_graph;
    int t0 = 0;
    if (graph && !B.two_steps) {   // B.cur is 0 here (decode_seed) and is 0 again after two steps
        // the first two tokens eagerly: every kernel's one-time function attributes are set outside the capture
        for (int k = 0; k < 2; ++k) {
            decode_one(B);
            launch_decode_advance(B.pred, b, cfg.eos, B.out, B.lead, B.flag, s);
        }
        t0 = 2;
        cudaGraph_t g = nullptr;
        CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
            for (int k = 0; k < 2; ++k) {
                decode_one(B);
                launch_decode_advance(B.pred, b, cfg.eos, B.out, B.lead, B.flag, s);
            }
        } catch (...) {
            cudaStreamEndCapture(s, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        CUDA_CHECK(cudaStreamEndCapture(s, &g));
        CUDA_CHECK(cudaGraphInstantiate(&B.two_steps, g, 0));
        CUDA_CHECK(cudaGraphDestroy(g));
    }
    for (int t = t0; t < steps; t += graph ? 2 : 1) {
        if (graph) {
            CUDA_CHECK(cudaGraphLaunch(B.two_steps, s));   // a step past the budget is dropped by k_decode_advance
        } else {
            decode_one(B);
            launch_decode_advance(B.pred, b, cfg.eos, B.out, B.lead, B.flag, s);
        }
        const int done = t + (graph ? 2 : 1);
        if ((done & 15) == 0 || done >= steps) {
            copy_sync(flag, B.flag, sizeof(flag), cudaMemcpyDeviceToHost);
            if (flag[0]) break;
        }
    }
    if (prof) {
        const double issue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
        CUDA_CHECK(cudaEventRecord(e1, s));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float dev_ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&dev_ms, e0, e1));
        fprintf(stderr, "[decode_prof] b=%d steps=%d device %.3f ms, host issue loop %.3f ms\n", b, flag[1], dev_ms, issue_ms);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    const int T = flag[1];
    std::vector<int32_t> tb((size_t)std::max(T, 1) * b);
    if (T > 0) copy_sync(tb.data(), B.out, sizeof(int) * (size_t)T * b, cudaMemcpyDeviceToHost);
    for (int i = 0; i < b; ++i)
        for (int t = 0; t < T; ++t) tokens[(size_t)i * steps + t] = tb[(size_t)t * b + i];
    CUDA_CHECK(cudaStreamSynchronize(s));
    return T;
}

// ------------------------------------------------------------------------------------------
// checkpoint container (tf.train.Saver stand-in, train.py:92-96,121): params + Adam slots + step
// ------------------------------------------------------------------------------------------
void Engine::save(const char* path) {
    FILE* f = fopen(path, "wb");
    if (!f) throw std::runtime_error(std::string("cannot open for writing: ") + path);
    std::vector<float> hp(nflat), hm(nflat), hv(nflat);
    copy_sync(hp.data(), p, nflat * sizeof(float), cudaMemcpyDeviceToHost);
    copy_sync(hm.data(), m, nflat * sizeof(float), cudaMemcpyDeviceToHost);
    copy_sync(hv.data(), v, nflat * sizeof(float), cudaMemcpyDeviceToHost);
    const char magic[8] = {'A', 'R', 'G', 'S', 'I', 'M', '0', '1'};
    fwrite(magic, 1, 8, f);
    int64_t hdr[2] = {step, (int64_t)params.size()};
    fwrite(hdr, sizeof(int64_t), 2, f);
    for (const ParamInfo& pi : params) {
        int32_t nl = (int32_t)pi.name.size();
        fwrite(&nl, 4, 1, f);
        fwrite(pi.name.data(), 1, nl, f);
        int64_t meta[3] = {pi.rank, pi.shape[0], pi.shape[1]};
        fwrite(meta, sizeof(int64_t), 3, f);
        fwrite(hp.data() + pi.off, sizeof(float), pi.n, f);
        fwrite(hm.data() + pi.off, sizeof(float), pi.n, f);
        fwrite(hv.data() + pi.off, sizeof(float), pi.n, f);
    }
    if (fclose(f) != 0) throw std::runtime_error("write failed");
}

void Engine::load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open: ") + path);
    char magic[8];
    int64_t hdr[2];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "ARGSIM01", 8) || fread(hdr, sizeof(int64_t), 2, f) != 2) {
        fclose(f);
        throw std::runtime_error("not an argsim_b200 checkpoint");
    }
    std::vector<float> hp(nflat, 0.f), hm(nflat, 0.f), hv(nflat, 0.f);
    for (int64_t k = 0; k < hdr[1]; ++k) {
        int32_t nl;
        if (fread(&nl, 4, 1, f) != 1 || nl <= 0 || nl > 4096) { fclose(f); throw std::runtime_error("corrupt checkpoint"); }
        std::string name(nl, 0);
        int64_t meta[3];
        if (fread(&name[0], 1, nl, f) != (size_t)nl || fread(meta, sizeof(int64_t), 3, f) != 3) { fclose(f); throw std::runtime_error("corrupt checkpoint"); }
        auto it = pindex.find(name);
        if (it == pindex.end()) { fclose(f); throw std::runtime_error("checkpoint has unknown tensor " + name); }
        const ParamInfo& pi = params[it->second];
        if (meta[0] != pi.rank || meta[1] != pi.shape[0] || meta[2] != pi.shape[1]) { fclose(f); throw std::runtime_error("shape mismatch for " + name); }
        if (fread(hp.data() + pi.off, sizeof(float), pi.n, f) != pi.n || fread(hm.data() + pi.off, sizeof(float), pi.n, f) != pi.n ||
            fread(hv.data() + pi.off, sizeof(float), pi.n, f) != pi.n) { fclose(f); throw std::runtime_error("truncated checkpoint"); }
    }
    fclose(f);
    copy_sync(p, hp.data(), nflat * sizeof(float), cudaMemcpyHostToDevice);
    copy_sync(m, hm.data(), nflat * sizeof(float), cudaMemcpyHostToDevice);
    copy_sync(v, hv.data(), nflat * sizeof(float), cudaMemcpyHostToDevice);
    refresh_shadow(0, nflat);
    CUDA_CHECK(cudaStreamSynchronize(st[0]));
    step = hdr[0];
}
