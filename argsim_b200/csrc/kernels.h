// kernels.h -- launch wrappers of the bandwidth-bound kernels (kernels.cu), the GEMMs
// (gemm_simt.cu, gemm_tc.cu) and the GRU recurrences (gru.cu).
#pragma once
#include "common.cuh"

// ---- kernels.cu -------------------------------------------------------------------------
// out[i,:] = table[ids[i],:]   (A5: tf.gather, model.py:111-112)
void launch_embed_gather_f32(const int* ids, long long n, const float* table, int D, float* out, cudaStream_t s);
void launch_embed_gather_bf16(const int* ids, long long n, const bf16* table, int D, bf16* out, cudaStream_t s);
// table_grad[ids[i],:] += dx[i,:]   (A18: IndexedSlices part of dE)
void launch_embed_scatter_add(const int* ids, long long n, const float* dx, int ld, int D, float* table_grad, cudaStream_t s);
// generic row gather: out[dst[i] or i, :] = in[src[i] or i, :]; reads fp32 or bf16, writes fp32 and/or bf16
void launch_row_gather(const float* in_f, const bf16* in_h, int ld_in, const int* src_idx, float* out_f, bf16* out_h,
                       int ld_out, const int* dst_idx, int n, int cols, cudaStream_t s);
// A9/A15: z = mu (+ exp(lv/2)*eps), kld_samp, sum(kld) -> stats[2].  eps is injected (eps_in) or Philox keyed by the GLOBAL row
// index of every row: row_ids[r] (device, b ints) or row0 + r
void launch_latent_fwd(const float* mulv, const float* eps_in, int b, int R, int train, uint64_t seed, uint64_t step,
                       long long row0, const int* row_ids, float* eps_used, float* z_f, bf16* z_h, float* kld_samp, double* stats, cudaStream_t s);
// A18: dmu = dz + a*mu ; dlv = dz*.5*exp(lv/2)*eps + a*.5*(exp(lv)-1)   (a = anneal/(b_global*R))
void launch_latent_bwd(const float* dz, const float* mulv, const float* eps, int b, int R, int train, float a,
                       float* dmulv_f, bf16* dmulv_h, cudaStream_t s);
// A12-A14 + A18: fused softmax cross entropy; logits overwritten by (softmax-onehot)*gscale when write_grad
void launch_ce_f32(float* logits, int ld, const int* labels, long long n, int V, float gscale, int write_grad,
                   float* loss_samp, float* err_samp, int* pred, double* stats, cudaStream_t s);
void launch_ce_bf16(bf16* logits, int ld, const int* labels, long long n, int V, float gscale, int write_grad,
                    float* loss_samp, float* err_samp, int* pred, double* stats, cudaStream_t s);
// A17: TF-form Adam over a flat buffer; optional bf16 shadow
void launch_adam(float* p, const float* g, float* m, float* v, bf16* shadow, long long n, float lr_t,
                 float b1, float b2, float eps, cudaStream_t s);
// out[c] = sum_r in[r,c]
void launch_colsum_f32(const float* in, int ld, long long rows, int cols, float* out, cudaStream_t s, int accumulate = 0);
void launch_colsum_bf16(const bf16* in, int ld, long long rows, int cols, float* out, cudaStream_t s, int accumulate = 0);
void launch_cast_bf16(const float* in, bf16* out, long long n, cudaStream_t s);
void launch_cast_f32(const bf16* in, float* out, long long n, cudaStream_t s);
// dst[dst_idx[i],:] += (or =) src[i,:]
void launch_row_scatter(const float* in, int ld_in, float* out, int ld_out, const int* dst_idx, int n, int cols,
                        int accumulate, cudaStream_t s);
void launch_fill(float* p, long long n, float v, cudaStream_t s);
void launch_fill_i32(int* p, long long n, int v, cudaStream_t s);
void launch_decode_advance(const int* pred, int b, int eos, int* out, int* lead, int* flag, cudaStream_t s);   // flag: stopped, steps kept, budget
// one GRU cell step on nb rows (decode(), model.py:204-219): state updated in place
void gru_generic_cell(const float* gx, int ld_gx, const float* R, const float* bR, float* state, float* gh_work, int nb,
                      int H, cudaStream_t s);

// ---- gemm_simt.cu / gemm_tc.cu ----------------------------------------------------------
// C(M,N) = alpha * A(M,K) * B(N,K)^T + bias[N] (+ C when accumulate).
// a_mn/b_mn = 0: operand stored (M|N rows, K cols) K-contiguous; 1: stored (K rows, M|N cols).
void gemm_simt(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, float* C, int ldc,
               int M, int N, int K, float alpha, const float* bias, int accumulate, bf16* C_h, cudaStream_t s);
// tcgen05 + TMA path (bf16 operands, fp32 accumulate in TMEM). Cf and/or Ch may be null.
// short_units > 0 (fp32 output): for GEMMs that share the chip with higher-priority kernels. The K range is split so
// that a work unit has at most short_units k-blocks and every unit is its own CTA (not persistent): an SM is held
// for a few microseconds at a time and a waiting block of the other kernel gets it between two units.
void gemm_tc_pdl(bool on);   // the calling thread's next k_gemm_tc2 launches use programmatic dependent launch
void gemm_tc(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, float* Cf, bf16* Ch, int ldc,
             int M, int N, int K, float alpha, const float* bias, int accumulate, cudaStream_t s, int short_units = 0);
void gemm_tc_init(int device);
void tma_encode_slice_rows_bf16(void* map_out /* CUtensorMap, 128 bytes */, const bf16* base, int ld, long long rows, int ns, int box_rows);
bool gemm_tc_available();

// ---- gru_generic.cu / gru_mma.cu ---------------------------------------------------------
// One direction-layer of a GRU in cuDNN form (gate order r,u,n; src/model.py:15,118-122,160)
// over a packed, length-sorted sequence set (plan.h).
struct GruFwdArgs {
    const float* gx;      // (rows, ld_gx) fp32: W x + bW for this direction (column offset applied)
    int ld_gx;
    const float* R_f;     // (3H,H) fp32 master
    const bf16* R_h;      // (3H,H) bf16 shadow (bf16 mode) or null
    const float* bR;      // (3H)
    const float* h0;      // (b,H) sorted order, or null = zeros.  Forward directions only.
    float* hs_f;          // (rows, ld_hs) outputs (column offset applied); either view may be null
    bf16* hs_h;
    int ld_hs;
    float* cache;         // (rows, 4H): r,u,n,q per row (training) or null
    int reverse;          // 1: runs t = Tmax-1 .. 0 (== tf.reverse_sequence o GRU o tf.reverse_sequence)
    float* hT = nullptr;  // persistent kernels, time-segmented launches: (b,H) state after the segment, or null
};
struct GruBwdArgs {
    const float* dhs;     // (rows, ld_dhs) gradient wrt outputs (column offset applied)
    int ld_dhs;
    const float* hs_f;    // forward outputs (one of the two views)
    const bf16* hs_h;
    int ld_hs;
    const float* h0;      // or null
    const float* cache;   // (rows, 4H)
    const float* R_f;
    const bf16* R_h;
    float* dgx_f;         // (rows, ld_dg) out: [dr,du,dn]      = grad wrt (W x + bW)
    bf16* dgx_h;
    float* dgh_f;         // (rows, ld_dg) out: [dr,du,dn*r]    = grad wrt (R h + bR)
    bf16* dgh_h;
    int ld_dg;
    float* hp_f;          // (rows, ld_hp) out: h_{prev} of every row (operand of the R wgrad)
    bf16* hp_h;
    int ld_hp;
    float* dh0;           // (b,H) += gradient wrt h0 (sorted order), or null
    int reverse;
    const float* dh_in = nullptr;   // time-segmented launches: gradient wrt the state after the segment (from the later one)
    float* dh_out = nullptr;        // ... and wrt the state before it (overwritten; handed to the earlier segment)
};
struct SeqPlan;
// generic path: one GEMM + one gate kernel per time step, fp32 (FP32_VALIDATE mode, any H)
void gru_generic_fwd(const GruFwdArgs* dirs, int ndir, const SeqPlan& P, int H, float* work, cudaStream_t* streams);
void gru_generic_bwd(const GruBwdArgs* dirs, int ndir, const SeqPlan& P, int H, float* work, cudaStream_t* streams);
size_t gru_generic_work_floats(int b, int H);   // per direction
// persistent path (H == 512, bf16 mode): register-stationary weights, LL exchange through L2
struct GruMmaCtx;
GruMmaCtx* gru_mma_create(int device);
void gru_mma_destroy(GruMmaCtx*);
bool gru_mma_supported(int H);
bool gru_mma_fits(const GruMmaCtx*, int ndir, int b);   // batch small enough for the resident-state kernel
// steps [t0, t0+Tseg) (Tseg < 0: the whole plan); launches that may overlap in time need different slots (0..3)
void gru_mma_fwd(GruMmaCtx*, const GruFwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact,
                 int H, cudaStream_t s, int t0 = 0, int Tseg = -1, int slot = 0, int alone = 0, int pad = 0);
// chunk: rows per chunk of this launch, 0 = the context's default (16, or 8 with ARGSIM_GRU_CHUNK=8), 8, 16
void gru_mma_bwd(GruMmaCtx*, const GruBwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact,
                 int H, cudaStream_t s, int t0 = 0, int Tseg = -1, int slot = 0, int alone = 0, int pad = 0, int chunk = 0);
// tensor-memory path (gru_tc.cu): the same recurrence on tcgen05, R_own resident in TMEM (TS form), N = 16..128 rows
// per MMA.  rows_per_slice: 0 = as many 16-row slices as groups fit, else the wanted slice size
struct GruTcCtx;
GruTcCtx* gru_tc_create(int device);
void gru_tc_destroy(GruTcCtx*);
bool gru_tc_supported(int H);
bool gru_tc_fits(const GruTcCtx*, int ndir, int b);
bool gru_tc_throughput(const GruTcCtx*, int ndir, int b);   // rows per slice large enough for the TMA-fed forward kernel
void gru_tc_fwd(GruTcCtx*, const GruFwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact,
                int H, cudaStream_t s, int t0 = 0, int Tseg = -1, int slot = 0, int rows_per_slice = 0, int pad = 0);
// one fused time step (TMA ring + tcgen05 + gate epilogue) for hidden sizes without a persistent kernel (gru_tc.cu)
bool gru_step_supported(int H, int b);
void gru_step_fwd(const GruFwdArgs* dirs, int ndir, const int* na, const long long* row0, int b, int H, float* const* state_f,
                  bf16* const* state_h_cur, bf16* const* state_h_next, cudaStream_t s, bool gx_fresh = false);
void gru_tc_bwd(GruTcCtx*, const GruBwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact,
                int H, cudaStream_t s, int t0 = 0, int Tseg = -1, int slot = 0, int rows_per_slice = 0, int pad = 0);
// returns the cycles from the first MMA issue to the completion of the last (K/16 MMAs round robin over nacc accumulators)
long long gru_tc_test_mma(const bf16* A, const bf16* B, float* D, int N, int K, int nacc, cudaStream_t s);
// xbench.cu: exchange-latency measurement hook
int xbench_run(int device, int method, int groups, int rows, int iters, double* cycles_per_iter, int* max_clusters);

// attentive=true (src/model.py:18-45,136-145): single-query multi-head attention over the packed encoder rows + residual
// layer norm; see kernels.cu
void launch_attn_fwd(const float* q, const float* K, const float* V, int b, int dim, int heads, const int* off, int Tmax,
                     const int* last_row, float* prob, float* y_f, bf16* y_h, cudaStream_t s);
void launch_attn_bwd(const float* dy, const float* q, const float* K, const float* V, const float* prob, int b, int dim, int heads,
                     const int* off, int Tmax, const int* last_row, float* dq_f, bf16* dq_h, float* dK_f, bf16* dK_h, float* dV_f,
                     bf16* dV_h, cudaStream_t s);
void launch_resid_ln_fwd(const float* h, const float* pp, const float* gamma, const float* beta, int b, int dim, float* xhat, float* rstd,
                         float* out_f, bf16* out_h, cudaStream_t s);
void launch_ln_bwd(const float* dout, const float* xhat, const float* rstd, const float* gamma, int b, int dim, float* dx_f, bf16* dx_h,
                   float* dgam_rows, cudaStream_t s);
