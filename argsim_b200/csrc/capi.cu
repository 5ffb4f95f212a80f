// capi.cu -- extern "C" shell of libargsim_b200.so (include/argsim_b200.h).  Plain pointers and
// sizes only; every entry point catches C++ exceptions and turns them into a status + message.
#include "engine.h"
#include "nccl_dyn.h"
#include <cuda_profiler_api.h>
#include <string.h>
#include <mutex>

struct argsim_handle {
    Engine* e = nullptr;
    std::string err;
    std::vector<const char*> tname_ptrs;
};

static thread_local std::string g_create_err;

#define API_BEGIN_NODRAIN(h)                           \
    if (!(h) || !(h)->e) return -1;                    \
    try {
// every entry point except the pipelined submit / wait pair first lets the steps in flight finish on the device
#define API_BEGIN(h)                                   \
    if (!(h) || !(h)->e) return -1;                    \
    try {                                              \
        (h)->e->drain();
#define API_END(h)                                     \
    } catch (const std::exception& ex) {               \
        (h)->err = ex.what();                          \
        cudaGetLastError();                            \
        return -2;                                     \
    }                                                  \
    return 0;

extern "C" {

const char* argsim_version(void) { return "argsim_b200 0.1 (sm_100a)"; }

const char* argsim_last_error(argsim_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int argsim_nccl_unique_id(uint8_t out[128]) {
    try {
        NcclApi& n = NcclApi::get();
        NcclApi::unique_id id;
        n.check(n.GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out, id.internal, 128);
    } catch (const std::exception& ex) {
        g_create_err = ex.what();
        return -2;
    }
    return 0;
}

int argsim_create(const argsim_config* cfg, argsim_handle** out) {
    if (!cfg || !out) return -1;
    try {
        argsim_handle* h = new argsim_handle();
        try {
            h->e = new Engine(*cfg);
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
    } catch (const std::exception& ex) {
        g_create_err = ex.what();
        cudaGetLastError();
        return -2;
    }
    return 0;
}

void argsim_destroy(argsim_handle* h) {
    if (!h) return;
    delete h->e;
    delete h;
}

int argsim_init_params(argsim_handle* h, uint64_t seed) {
    API_BEGIN(h) h->e->init_params(seed);
    API_END(h)
}
int argsim_param_count(argsim_handle* h, int32_t* n) {
    API_BEGIN(h) *n = (int32_t)h->e->params.size();
    API_END(h)
}
int argsim_param_info(argsim_handle* h, int32_t i, const char** name, int32_t* rank, int64_t shape[2]) {
    API_BEGIN(h)
    if (i < 0 || i >= (int)h->e->params.size()) throw std::runtime_error("parameter index out of range");
    const ParamInfo& pi = h->e->params[i];
    if (name) *name = pi.name.c_str();
    if (rank) *rank = pi.rank;
    if (shape) { shape[0] = pi.shape[0]; shape[1] = pi.shape[1]; }
    API_END(h)
}
int argsim_get_param(argsim_handle* h, const char* name, float* dst) {
    API_BEGIN(h) h->e->get_flat(h->e->p, name, dst);
    API_END(h)
}
int argsim_set_param(argsim_handle* h, const char* name, const float* src) {
    API_BEGIN(h) h->e->set_param(name, src);
    API_END(h)
}
int argsim_get_grad(argsim_handle* h, const char* name, float* dst) {
    API_BEGIN(h) h->e->get_flat(h->e->g, name, dst);
    API_END(h)
}
int argsim_get_opt_state(argsim_handle* h, const char* name, float* m, float* v) {
    API_BEGIN(h)
    if (m) h->e->get_flat(h->e->m, name, m);
    if (v) h->e->get_flat(h->e->v, name, v);
    API_END(h)
}
int argsim_set_opt_state(argsim_handle* h, const char* name, const float* m, const float* v) {
    API_BEGIN(h)
    if (m) h->e->set_flat(h->e->m, name, m);
    if (v) h->e->set_flat(h->e->v, name, v);
    API_END(h)
}
int argsim_get_step(argsim_handle* h, int64_t* step) {
    API_BEGIN(h) *step = h->e->step;
    API_END(h)
}
int argsim_set_step(argsim_handle* h, int64_t step) {
    API_BEGIN(h) h->e->step = step;
    API_END(h)
}
int argsim_set_seed(argsim_handle* h, uint64_t seed) {
    API_BEGIN(h) h->e->seed = seed;
    API_END(h)
}

int argsim_train_step(argsim_handle* h, const int32_t* src, const int32_t* tgt, int32_t b, int32_t T_src, int32_t T_tgt,
                      const uint8_t* keep_mask, const float* eps, int64_t n_tokens_global, int64_t b_global,
                      int64_t row0_global, argsim_step_stats* out) {
    API_BEGIN(h)
    h->e->train_step(src, tgt, b, T_src, T_tgt, keep_mask, eps, n_tokens_global, b_global, row0_global, true, out);
    API_END(h)
}
int argsim_train_step_submit(argsim_handle* h, const int32_t* src, const int32_t* tgt, int32_t b, int32_t T_src, int32_t T_tgt,
                             const uint8_t* keep_mask, const float* eps, int64_t n_tokens_global, int64_t b_global,
                             int64_t row0_global) {
    API_BEGIN_NODRAIN(h)
    h->e->train_step_submit(src, tgt, b, T_src, T_tgt, keep_mask, eps, n_tokens_global, b_global, row0_global, true);
    API_END(h)
}
int argsim_set_global_rows(argsim_handle* h, const int64_t* rows, int32_t b) {
    API_BEGIN_NODRAIN(h)
    h->e->set_global_rows(rows, b);
    API_END(h)
}
int argsim_train_step_wait(argsim_handle* h, argsim_step_stats* out) {
    API_BEGIN_NODRAIN(h)
    h->e->train_step_wait(out);
    API_END(h)
}
int argsim_grad_step(argsim_handle* h, const int32_t* src, const int32_t* tgt, int32_t b, int32_t T_src, int32_t T_tgt,
                     const uint8_t* keep_mask, const float* eps, int64_t n_tokens_global, int64_t b_global,
                     int64_t row0_global, argsim_step_stats* out) {
    API_BEGIN(h)
    h->e->train_step(src, tgt, b, T_src, T_tgt, keep_mask, eps, n_tokens_global, b_global, row0_global, false, out);
    API_END(h)
}
int argsim_eval_step(argsim_handle* h, const int32_t* src, const int32_t* tgt, int32_t b, int32_t T_src, int32_t T_tgt,
                     float* errt_samp, float* loss_gen_samp, int64_t cap_rows, float* loss_kld_samp, int64_t* n_rows,
                     int32_t* pred_or_null) {
    API_BEGIN(h)
    h->e->eval_step(src, tgt, b, T_src, T_tgt, errt_samp, loss_gen_samp, cap_rows, loss_kld_samp, n_rows, pred_or_null);
    API_END(h)
}
int argsim_embed(argsim_handle* h, const int32_t* src, int32_t b, int32_t T, float* mu_out) {
    API_BEGIN(h) h->e->embed(src, b, T, mu_out);
    API_END(h)
}
int argsim_decode_init(argsim_handle* h, const float* z, int32_t b, float* state) {
    API_BEGIN(h) h->e->decode_init(z, b, state);
    API_END(h)
}
int argsim_decode_step(argsim_handle* h, const int32_t* lead, int32_t b, float* state_inout, int32_t* pred) {
    API_BEGIN(h) h->e->decode_step(lead, b, state_inout, pred);
    API_END(h)
}
int argsim_decode(argsim_handle* h, const float* z, int32_t b, int32_t steps, int32_t* tokens, int32_t* t_out) {
    API_BEGIN(h)
    const int T = h->e->decode_loop(z, b, steps, tokens);
    if (t_out) *t_out = T;
    API_END(h)
}
int argsim_save(argsim_handle* h, const char* path) {
    API_BEGIN(h) h->e->save(path);
    API_END(h)
}
int argsim_load(argsim_handle* h, const char* path) {
    API_BEGIN(h) h->e->load(path);
    API_END(h)
}
int argsim_bench_resident(argsim_handle* h, int32_t iters, float* ms_per_step) {
    API_BEGIN(h) h->e->bench_resident(iters, ms_per_step);
    API_END(h)
}
int argsim_launch_count(argsim_handle* h, int64_t* n) {
    API_BEGIN(h) *n = g_launch_count;
    API_END(h)
}
int argsim_profiler(argsim_handle* h, int32_t on) {
    API_BEGIN(h)
    CUDA_CHECK(cudaDeviceSynchronize());
    if (on) {
        h->e->nvtx = true;
        CUDA_CHECK(cudaProfilerStart());
    } else {
        CUDA_CHECK(cudaProfilerStop());
        h->e->nvtx = false;
    }
    API_END(h)
}
int argsim_last_timings(argsim_handle* h, int32_t cap, const char** names, float* ms) {
    if (!h || !h->e) return -1;
    int n = (int)std::min<size_t>(h->e->tnames.size(), (size_t)std::max(cap, 0));
    for (int i = 0; i < n; ++i) {
        names[i] = h->e->tnames[i].c_str();
        ms[i] = h->e->tms[i];
    }
    return n;
}

int argsim_test_gemm(int32_t device, int32_t impl, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                     const float* A, const float* B, const float* bias, float alpha, int32_t accumulate, float* C,
                     float* ms_out) {
    try {
        CUDA_CHECK(cudaSetDevice(device));
        const size_t na = (size_t)M * K, nb = (size_t)N * K, nc = (size_t)M * N;
        float *dA, *dB, *dC, *dbias = nullptr;
        CUDA_CHECK(cudaMalloc(&dA, na * 4)); CUDA_CHECK(cudaMalloc(&dB, nb * 4)); CUDA_CHECK(cudaMalloc(&dC, nc * 4));
        CUDA_CHECK(cudaMemcpy(dA, A, na * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(dB, B, nb * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(dC, C, nc * 4, cudaMemcpyHostToDevice));
        if (bias) {
            CUDA_CHECK(cudaMalloc(&dbias, (size_t)N * 4));
            CUDA_CHECK(cudaMemcpy(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
        }
        const int lda = a_mn ? M : K, ldb = b_mn ? N : K;
        cudaEvent_t e0, e1;
        CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
        if (impl == 0) {
            CUDA_CHECK(cudaEventRecord(e0, 0));
            gemm_simt(dA, lda, a_mn, dB, ldb, b_mn, dC, N, M, N, K, alpha, dbias, accumulate, nullptr, 0);
            CUDA_CHECK(cudaEventRecord(e1, 0));
        } else {
            gemm_tc_init(device);
            if (!gemm_tc_available()) throw std::runtime_error("tcgen05 GEMM path unavailable on this device");
            bf16 *hA, *hB;
            CUDA_CHECK(cudaMalloc(&hA, na * 2)); CUDA_CHECK(cudaMalloc(&hB, nb * 2));
            launch_cast_bf16(dA, hA, (long long)na, 0);
            launch_cast_bf16(dB, hB, (long long)nb, 0);
            bf16* hC = nullptr;
            if (impl == 2) CUDA_CHECK(cudaMalloc(&hC, nc * 2));   // bf16 output (activation-typed C)
            CUDA_CHECK(cudaEventRecord(e0, 0));
            gemm_tc(hA, lda, a_mn, hB, ldb, b_mn, impl == 2 ? nullptr : dC, hC, N, M, N, K, alpha, dbias, impl == 2 ? 0 : accumulate, 0);
            CUDA_CHECK(cudaEventRecord(e1, 0));
            if (hC) launch_cast_f32(hC, dC, (long long)nc, 0);
            CUDA_CHECK(cudaDeviceSynchronize());
            cudaFree(hA); cudaFree(hB); cudaFree(hC);
        }
        CUDA_CHECK(cudaDeviceSynchronize());
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_out) *ms_out = ms;
        CUDA_CHECK(cudaMemcpy(C, dC, nc * 4, cudaMemcpyDeviceToHost));
        cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbias);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    } catch (const std::exception& ex) {
        g_create_err = ex.what();
        cudaGetLastError();
        return -2;
    }
    return 0;
}

int argsim_test_ts_mma(int32_t device, int32_t N, int32_t K, int32_t nacc, const float* A, const float* B, float* D, int64_t* cycles) {
    try {
        CUDA_CHECK(cudaSetDevice(device));
        const size_t na = (size_t)128 * K, nb = (size_t)N * K, nd = (size_t)128 * N;
        float *dA, *dB, *dD;
        bf16 *hA, *hB;
        CUDA_CHECK(cudaMalloc(&dA, na * 4)); CUDA_CHECK(cudaMalloc(&dB, nb * 4)); CUDA_CHECK(cudaMalloc(&dD, nd * 4));
        CUDA_CHECK(cudaMalloc(&hA, na * 2)); CUDA_CHECK(cudaMalloc(&hB, nb * 2));
        CUDA_CHECK(cudaMemcpy(dA, A, na * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(dB, B, nb * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemset(dD, 0, nd * 4));
        launch_cast_bf16(dA, hA, (long long)na, 0);
        launch_cast_bf16(dB, hB, (long long)nb, 0);
        const long long cyc = gru_tc_test_mma(hA, hB, dD, N, K, nacc, 0);
        if (cycles) *cycles = cyc;
        CUDA_CHECK(cudaDeviceSynchronize());
        CUDA_CHECK(cudaMemcpy(D, dD, nd * 4, cudaMemcpyDeviceToHost));
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(hA); cudaFree(hB);
    } catch (const std::exception& ex) {
        g_create_err = ex.what();
        cudaGetLastError();
        return -2;
    }
    return 0;
}

int argsim_test_softmax_ce(int32_t device, int32_t bf16_mode, int64_t n, int32_t V, const float* logits, const int32_t* labels,
                           float gscale, int32_t write_grad, float* grad_out, float* loss_samp, float* err_samp, int32_t* pred,
                           double stats[2]) {
    try {
        CUDA_CHECK(cudaSetDevice(device));
        const size_t nl = (size_t)n * V;
        float *dl, *dls, *der;
        int *dlab = nullptr, *dpred;
        double* dst;
        bf16* dh = nullptr;
        CUDA_CHECK(cudaMalloc(&dl, std::max<size_t>(nl, 1) * 4));
        CUDA_CHECK(cudaMalloc(&dls, std::max<size_t>(n, 1) * 4)); CUDA_CHECK(cudaMalloc(&der, std::max<size_t>(n, 1) * 4));
        CUDA_CHECK(cudaMalloc(&dpred, std::max<size_t>(n, 1) * 4)); CUDA_CHECK(cudaMalloc(&dst, 16));
        CUDA_CHECK(cudaMemset(dst, 0, 16));
        CUDA_CHECK(cudaMemcpy(dl, logits, nl * 4, cudaMemcpyHostToDevice));
        if (labels) {
            CUDA_CHECK(cudaMalloc(&dlab, std::max<size_t>(n, 1) * 4));
            CUDA_CHECK(cudaMemcpy(dlab, labels, (size_t)n * 4, cudaMemcpyHostToDevice));
        }
        if (bf16_mode) {
            CUDA_CHECK(cudaMalloc(&dh, std::max<size_t>(nl, 1) * 2));
            launch_cast_bf16(dl, dh, (long long)nl, 0);
            launch_ce_bf16(dh, V, dlab, n, V, gscale, write_grad, dls, der, dpred, dst, 0);
            launch_cast_f32(dh, dl, (long long)nl, 0);
        } else {
            launch_ce_f32(dl, V, dlab, n, V, gscale, write_grad, dls, der, dpred, dst, 0);
        }
        CUDA_CHECK(cudaDeviceSynchronize());
        if (grad_out) CUDA_CHECK(cudaMemcpy(grad_out, dl, nl * 4, cudaMemcpyDeviceToHost));
        if (loss_samp) CUDA_CHECK(cudaMemcpy(loss_samp, dls, (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (err_samp) CUDA_CHECK(cudaMemcpy(err_samp, der, (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (pred) CUDA_CHECK(cudaMemcpy(pred, dpred, (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (stats) CUDA_CHECK(cudaMemcpy(stats, dst, 16, cudaMemcpyDeviceToHost));
        cudaFree(dl); cudaFree(dls); cudaFree(der); cudaFree(dpred); cudaFree(dst); cudaFree(dlab); cudaFree(dh);
    } catch (const std::exception& ex) {
        g_create_err = ex.what();
        cudaGetLastError();
        return -2;
    }
    return 0;
}

int argsim_bench_kernel(argsim_handle* h, const char* which, int64_t rows, int32_t iters, float* ms, double* algo_bytes,
                        double* algo_flops) {
    API_BEGIN(h)
    Engine* e = h->e;
    const std::string w = which;
    const int V = e->V, D = e->D;
    cudaStream_t s = 0;
    // L2 flush buffer (> 126 MB), written between timed launches
    const size_t flush_n = (size_t)160 << 20;
    char* flush;
    CUDA_CHECK(cudaMalloc(&flush, flush_n));
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    float total = 0.f;
    double bytes = 0, flops = 0;
    if (w == "softmax_ce" || w == "logits_gemm") {
        bf16 *lg, *ho;
        int* lab;
        float *ls, *es; int* pr; double* stt;
        CUDA_CHECK(cudaMalloc(&lg, (size_t)rows * V * 2));
        CUDA_CHECK(cudaMalloc(&ho, (size_t)rows * D * 2));
        CUDA_CHECK(cudaMalloc(&lab, rows * 4)); CUDA_CHECK(cudaMalloc(&ls, rows * 4)); CUDA_CHECK(cudaMalloc(&es, rows * 4));
        CUDA_CHECK(cudaMalloc(&pr, rows * 4)); CUDA_CHECK(cudaMalloc(&stt, 32));
        CUDA_CHECK(cudaMemset(lab, 0, rows * 4)); CUDA_CHECK(cudaMemset(stt, 0, 32));
        CUDA_CHECK(cudaMemset(ho, 0, (size_t)rows * D * 2));
        const ParamInfo* emb = nullptr;
        for (const ParamInfo& pi : e->params) if (pi.name == "embed/embedding") emb = &pi;
        for (int it = 0; it < iters + 1; ++it) {
            CUDA_CHECK(cudaMemsetAsync(lg, 0x3c, (size_t)rows * V * 2, s));  // bf16 0x3c3c ~ 0.0115
            CUDA_CHECK(cudaMemsetAsync(flush, it, flush_n, s));
            CUDA_CHECK(cudaEventRecord(e0, s));
            if (w == "softmax_ce") {
                launch_ce_bf16(lg, V, lab, rows, V, 1.0f / (float)rows, 1, ls, es, pr, stt, s);
            } else {
                if (!e->ph) throw std::runtime_error("logits_gemm bench needs BF16 precision");
                gemm_tc(ho, D, 0, e->ph + emb->off, D, 0, nullptr, lg, V, (int)rows, V, D, 1.f, nullptr, 0, s);
            }
            CUDA_CHECK(cudaEventRecord(e1, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            float t; cudaEventElapsedTime(&t, e0, e1);
            if (it > 0) total += t;
        }
        bytes = (w == "softmax_ce") ? 4.0 * rows * V : 2.0 * rows * V + 2.0 * rows * D + 2.0 * V * D;
        flops = (w == "softmax_ce") ? 0 : 2.0 * rows * V * D;
        cudaFree(lg); cudaFree(ho); cudaFree(lab); cudaFree(ls); cudaFree(es); cudaFree(pr); cudaFree(stt);
    } else if (w == "adam") {
        const long long n = (long long)e->nflat;
        float *pp, *gg, *mm, *vv; bf16* sh = nullptr;
        CUDA_CHECK(cudaMalloc(&pp, n * 4)); CUDA_CHECK(cudaMalloc(&gg, n * 4)); CUDA_CHECK(cudaMalloc(&mm, n * 4)); CUDA_CHECK(cudaMalloc(&vv, n * 4));
        if (e->ph) CUDA_CHECK(cudaMalloc(&sh, n * 2));
        CUDA_CHECK(cudaMemset(pp, 0, n * 4)); CUDA_CHECK(cudaMemset(gg, 0, n * 4)); CUDA_CHECK(cudaMemset(mm, 0, n * 4)); CUDA_CHECK(cudaMemset(vv, 0, n * 4));
        for (int it = 0; it < iters + 1; ++it) {
            CUDA_CHECK(cudaMemsetAsync(flush, it, flush_n, s));
            CUDA_CHECK(cudaEventRecord(e0, s));
            launch_adam(pp, gg, mm, vv, sh, n, 1e-3f, 0.9f, 0.999f, 1e-8f, s);
            CUDA_CHECK(cudaEventRecord(e1, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            float t; cudaEventElapsedTime(&t, e0, e1);
            if (it > 0) total += t;
        }
        bytes = (double)n * (28.0 + (sh ? 2.0 : 0.0));
        cudaFree(pp); cudaFree(gg); cudaFree(mm); cudaFree(vv); cudaFree(sh);
    } else if (w == "embed_gather") {
        bf16 *tab, *out; int* ids;
        CUDA_CHECK(cudaMalloc(&tab, (size_t)V * D * 2)); CUDA_CHECK(cudaMalloc(&out, (size_t)rows * D * 2)); CUDA_CHECK(cudaMalloc(&ids, rows * 4));
        std::vector<int> hid(rows);
        for (int64_t i = 0; i < rows; ++i) hid[i] = (int)((i * 2654435761u) % (uint32_t)V);
        CUDA_CHECK(cudaMemcpy(ids, hid.data(), rows * 4, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemset(tab, 0, (size_t)V * D * 2));
        for (int it = 0; it < iters + 1; ++it) {
            CUDA_CHECK(cudaMemsetAsync(flush, it, flush_n, s));
            CUDA_CHECK(cudaEventRecord(e0, s));
            launch_embed_gather_bf16(ids, rows, tab, D, out, s);
            CUDA_CHECK(cudaEventRecord(e1, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            float t; cudaEventElapsedTime(&t, e0, e1);
            if (it > 0) total += t;
        }
        bytes = 4.0 * rows * D;
        cudaFree(tab); cudaFree(out); cudaFree(ids);
    } else {
        throw std::runtime_error("unknown kernel name: " + w);
    }
    if (ms) *ms = total / (float)std::max(iters, 1);
    if (algo_bytes) *algo_bytes = bytes;
    if (algo_flops) *algo_flops = flops;
    cudaFree(flush);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    API_END(h)
}

int argsim_plan_batch(const int32_t* src, const int32_t* tgt, int32_t b, int32_t T_src, int32_t T_tgt, int32_t bos,
                      int32_t eos, const uint8_t* keep, int32_t* len_src, int32_t* len_tgt, int64_t counts[4],
                      int32_t* ids_src, int32_t* lead, int32_t* gold, int32_t* ref_row, int32_t* enc_last,
                      int32_t* perm_src, int32_t* perm_dec, char* err, int32_t err_cap) {
    BatchPlan P;
    DropoutSpec drop;
    drop.train = keep ? 1 : 0;
    drop.keep = keep;
    std::string e;
    try {
        e = build_batch_plan(src, b, T_src, tgt, T_tgt, bos, eos, tgt ? 1 : 0, drop, &P);
    } catch (const std::exception& ex) {
        e = ex.what();
    }
    if (!e.empty()) {
        if (err && err_cap > 0) {
            strncpy(err, e.c_str(), err_cap - 1);
            err[err_cap - 1] = 0;
        }
        return -2;
    }
    auto cp = [](int32_t* dst, const std::vector<int>& v) {
        if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(int));
    };
    cp(len_src, P.len_src); cp(len_tgt, P.len_tgt); cp(ids_src, P.ids_src); cp(lead, P.ids_lead); cp(gold, P.labels);
    cp(ref_row, P.ref_row); cp(enc_last, P.enc_last); cp(perm_src, P.enc.perm); cp(perm_dec, P.dec.perm);
    if (counts) {
        counts[0] = P.enc.rows; counts[1] = P.dec.rows; counts[2] = P.enc.Tmax; counts[3] = P.dec.Tmax;
    }
    return 0;
}

void argsim_schedule(int64_t step, float accelerate, float learn_rate, float* keepwd, float* anneal, float* update) {
    schedule_f32(step, accelerate, learn_rate, keepwd, anneal, update);
}

}  // extern "C"
