/* argsim_b200.h -- C ABI of libargsim_b200.so
 *
 * Drop-in boundary for the ONE hot path of argsim/argsim: the sequence-VAE ELBO
 * training step (src/model.py:48-189 driven by src/train.py:115-121) and the
 * encoder-only mu embedding (src/model.py:194-201, src/eval_embed_reason.py:38).
 *
 * The reference has no FFI of its own: its boundary is `sess.run(fetches, feed_dict)`
 * on the TF graph built by `vAe(...)`.  Each entry point below replaces one fetch set
 * of that graph (cited per function); the Python facade argsim_b200/model.py binds
 * them through ctypes and re-exposes the reference's call surface.
 *
 * Conventions (inherited from the reference call sites):
 *   - token matrices are C-contiguous int32 (b,T), batch-major, eos-padded, every row
 *     has >= 1 non-eos token; outputs are float32 / int32 host buffers owned by the caller;
 *   - calls are blocking; a handle is not thread-safe;
 *   - every function returns 0 on success, <0 on error; argsim_last_error() returns a
 *     message owned by the library (valid until the next call on that handle; pass NULL
 *     for errors raised by argsim_create);
 *   - parameters are exchanged in fp32 in the canonical layout listed in
 *     oracle/vae_oracle.py (dense kernels are (in,out) like tf.layers.dense, GRU
 *     matrices are (3H,in)/(3H,H) in gate order r,u,n).
 *   - there is NO CPU fallback: without a CUDA device argsim_create fails.
 */
#ifndef ARGSIM_B200_H
#define ARGSIM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct argsim_handle argsim_handle;

enum { ARGSIM_FP32_VALIDATE = 0, ARGSIM_BF16 = 1 };

/* kwargs of vAe() (src/model.py:48-64) == config.json "model" section, plus runtime knobs */
typedef struct argsim_config {
    int32_t dim_tgt, dim_emb, dim_rep, rnn_layers;
    int32_t bidirectional, bidir_stacked, attentive, logit_use_embed;
                                  /* attentive (src/model.py:136-145): the final state attends over its sequence's outputs,
                                     8 heads, residual + layer norm; adds encode/cata/{q,k,v,p}/{kernel,bias} and
                                     encode/cata/LayerNorm/{gamma,beta}.  The reference branch is marked "todo fixme" and
                                     cannot execute; the repaired semantics are stated in DESIGN.md section 1, row A8 */
    float   accelerate, learn_rate;
    int32_t bos, eos;
    int32_t precision;            /* ARGSIM_FP32_VALIDATE | ARGSIM_BF16 */
    int32_t max_batch, max_len;   /* per-rank capacity: rows, tokens per row (without bos/eos pad) */
    int32_t device;               /* CUDA device ordinal */
    int32_t nranks, rank;         /* data parallel world (1,0 = single GPU) */
    uint8_t nccl_id[128];         /* ncclUniqueId from argsim_nccl_unique_id on rank 0 (nranks>1) */
    int32_t flags;                /* bit2 (4): BF16 mode with the generic per-step GRU instead of the persistent
                                     kernels; bit3 (8): record per-kernel CUDA-event timers (argsim_last_timings) */
} argsim_config;

/* per-step scalars; sums are over the GLOBAL batch when nranks>1 */
typedef struct argsim_step_stats {
    float   loss, loss_gen, loss_kld, errt;      /* model.py:185,181,184,177 */
    float   rate_keepwd, rate_anneal, rate_update; /* model.py:78-80, for the step just run */
    int64_t n_tokens;                            /* N = sum(len_tgt+1), global */
    int64_t step;                                /* global step AFTER the update */
} argsim_step_stats;

const char* argsim_version(void);
const char* argsim_last_error(argsim_handle*);
int  argsim_nccl_unique_id(uint8_t out[128]);                         /* rank 0, then broadcast by the host */

int  argsim_create(const argsim_config* cfg, argsim_handle** out);    /* vAe(...) graph build + Session, model.py:48, train.py:91 */
void argsim_destroy(argsim_handle*);
int  argsim_init_params(argsim_handle*, uint64_t seed);               /* global_variables_initializer, model.py:8-15,109 */

int  argsim_param_count(argsim_handle*, int32_t* n);
int  argsim_param_info(argsim_handle*, int32_t i, const char** name, int32_t* rank, int64_t shape[2]);
int  argsim_get_param(argsim_handle*, const char* name, float* host_dst);
int  argsim_set_param(argsim_handle*, const char* name, const float* host_src);
int  argsim_get_grad(argsim_handle*, const char* name, float* host_dst);   /* test only: d loss / d param of the last train step */
int  argsim_get_opt_state(argsim_handle*, const char* name, float* m, float* v);  /* Adam slots (Saver saves them, train.py:92) */
int  argsim_set_opt_state(argsim_handle*, const char* name, const float* m, const float* v);
int  argsim_get_step(argsim_handle*, int64_t* step);                  /* model.step, train.py:119 */
int  argsim_set_step(argsim_handle*, int64_t step);
int  argsim_set_seed(argsim_handle*, uint64_t seed);                  /* tf.set_random_seed, train.py:45 */

/* sess.run(model_train.train_step), train.py:118.  src/tgt: this rank's rows.
 * keep_mask (b,T_tgt) uint8 and eps (b,dim_rep) float may be NULL (device Philox RNG keyed by
 * seed, step, global row, position) or injected for parity runs (TF's streams cannot be matched).
 * n_tokens_global / b_global: normalisers of the two means (model.py:181,184) over the
 * GLOBAL batch; pass 0 to use this rank's own counts (single GPU).
 * row0_global: index of this rank's first row in the global batch (RNG keying). */
int  argsim_train_step(argsim_handle*, const int32_t* src, const int32_t* tgt, int32_t b,
                       int32_t T_src, int32_t T_tgt, const uint8_t* keep_mask, const float* eps,
                       int64_t n_tokens_global, int64_t b_global, int64_t row0_global,
                       argsim_step_stats* out);
/* Pipelined form of the same call.  src/train.py:118 fetches nothing but the op (sess.run(model.train_step)), so the
 * caller does not need the step to have finished: submit returns once the step is enqueued (host plan and H2D staging
 * done, src/tgt/keep_mask/eps no longer referenced) and wait returns the statistics of the OLDEST un-waited step.
 * At most two steps may be un-waited, i.e. the pattern is submit(n+1); wait(n): the host side of step n+1 overlaps the
 * device's step n.  Every other entry point first lets the steps in flight finish; argsim_train_step itself refuses
 * to run while un-waited steps exist. */
int  argsim_train_step_submit(argsim_handle*, const int32_t* src, const int32_t* tgt, int32_t b,
                              int32_t T_src, int32_t T_tgt, const uint8_t* keep_mask, const float* eps,
                              int64_t n_tokens_global, int64_t b_global, int64_t row0_global);
int  argsim_train_step_wait(argsim_handle*, argsim_step_stats* out);
/* Data parallel: global index of every row of the NEXT train / grad step (consumed by that call;
 * NULL or b = 0 clears).  The Philox streams of word dropout (model.py:94) and eps (model.py:150) are keyed by
 * (seed, step, global row, position), so with the rows' indices in the GLOBAL batch given here the un-injected
 * randomness of a step does not depend on the number of ranks or on how the rows were dealt; without it rows are
 * keyed row0_global + i. */
int  argsim_set_global_rows(argsim_handle*, const int64_t* rows, int32_t b);
/* same step but stops before Adam / step increment (gradient parity, test only) */
int  argsim_grad_step(argsim_handle*, const int32_t* src, const int32_t* tgt, int32_t b,
                      int32_t T_src, int32_t T_tgt, const uint8_t* keep_mask, const float* eps,
                      int64_t n_tokens_global, int64_t b_global, int64_t row0_global,
                      argsim_step_stats* out);

/* sess.run((errt_samp, loss_gen_samp, loss_kld_samp)) on the 'valid' graph, train.py:109-110.
 * Per-row outputs are in the reference's boolean_mask order (time-major over the original
 * batch order).  cap_rows = capacity of the two per-row buffers; *n_rows receives N. */
int  argsim_eval_step(argsim_handle*, const int32_t* src, const int32_t* tgt, int32_t b,
                      int32_t T_src, int32_t T_tgt, float* errt_samp, float* loss_gen_samp,
                      int64_t cap_rows, float* loss_kld_samp /* b*dim_rep */, int64_t* n_rows,
                      int32_t* pred_or_null /* cap_rows */);

/* model.z.eval({model.src: ...}) in 'infer'/'valid' mode == mu, eval_embed_reason.py:38; encode(), model.py:194-201 */
int  argsim_embed(argsim_handle*, const int32_t* src, int32_t b, int32_t T, float* mu_out /* b*dim_rep */);

/* decode(), model.py:204-219: state_in.eval({z}) then one (pred,state_ex) step per call */
int  argsim_decode_init(argsim_handle*, const float* z, int32_t b, float* state /* L*b*H */);
int  argsim_decode_step(argsim_handle*, const int32_t* lead /* b */, int32_t b,
                        float* state_inout /* L*b*H */, int32_t* pred /* b */);

/* the whole decode() loop of model.py:204-219 on the device: x = bos; repeat {pred, state} until
 * every row emits eos or `steps` steps; no host round trip per token.  tokens: (b, steps) int32 row-major, the
 * first *t_out columns are valid (the all-eos step is not kept, like the reference's break before append). */
int  argsim_decode(argsim_handle*, const float* z /* b*dim_rep */, int32_t b, int32_t steps,
                   int32_t* tokens /* b*steps */, int32_t* t_out);

/* tf.train.Saver save/restore (train.py:92-96,121): own flat container (params, Adam slots, step) */
int  argsim_save(argsim_handle*, const char* path);
int  argsim_load(argsim_handle*, const char* path);

/* ---- measurement / test hooks (not part of the reference surface) ------------------- */
/* device-resident replay of the last staged batch: runs `iters` full train steps (fwd+bwd+
 * [allreduce]+Adam) without host<->device copies and returns the mean device time per step. */
int  argsim_bench_resident(argsim_handle*, int32_t iters, float* ms_per_step);
/* number of kernels this library launched since creation (bench.py's gpu_launches) */
int  argsim_launch_count(argsim_handle*, int64_t* n);
/* --profile of src/train.py:76-82 / profile() of src/util_tf.py:9-13 (one traced sess.run): on = 1 calls
 * cudaProfilerStart() and makes every device program an NVTX range ("argsim:train_step" / "argsim:eval_step" /
 * "argsim:embed") with an NVTX mark at each phase boundary; on = 0 calls cudaProfilerStop().  For
 * `ncu / nsys --capture-range=cudaProfilerApi`. */
int  argsim_profiler(argsim_handle*, int32_t on);
/* per-phase device timings of the last step, name/ms pairs; returns count */
int  argsim_last_timings(argsim_handle*, int32_t cap, const char** names, float* ms);
/* unit-test hook for the GEMM kernels: C(M,N) = alpha * op(A) op(B)^T (+bias) ; layouts:
 * a_mn/b_mn = 0 operand stored (rows=M|N, cols=K) K-contiguous, 1 stored (K, M|N).  host fp32 in/out.
 * impl 0 = SIMT fp32, 1 = tcgen05 bf16 operands -> fp32 C, 2 = tcgen05 with a bf16 C (accumulate ignored). */
int  argsim_test_gemm(int32_t device, int32_t impl, int32_t M, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                      const float* A, const float* B, const float* bias_or_null, float alpha,
                      int32_t accumulate, float* C_inout, float* ms_or_null);
/* unit-test hook for the fused softmax cross-entropy kernel (model.py:170-181 + its gradient): host fp32 logits
 * (n,V) are copied to the device (rounded to bf16 when bf16_mode), the kernel runs once, and the in-place gradient
 * (softmax - onehot) * gscale, per-row loss / error flag / argmax and the two fp64 sums {loss, errors} come back.
 * labels may be NULL (argmax only); any output may be NULL. */
int  argsim_test_softmax_ce(int32_t device, int32_t bf16_mode, int64_t n, int32_t V, const float* logits,
                            const int32_t* labels_or_null, float gscale, int32_t write_grad, float* grad_out,
                            float* loss_samp, float* err_samp, int32_t* pred, double stats[2]);
/* unit-test hook for the building blocks of the tensor-memory recurrence (csrc/gru_tc.cu): D(128,N) = A(128,K) . B(N,K)^T
 * with A written to TMEM by tcgen05.st and read by the TS form of tcgen05.mma, B staged in 128-byte-swizzled shared
 * memory; operands rounded to bf16, fp32 out.  N in 16..128 (step 16), K in 64..512 (step 64).  The K/16 instructions
 * go round robin over `nacc` accumulators (summed on read-back); *cycles = first issue -> completion of the last. */
int  argsim_test_ts_mma(int32_t device, int32_t N, int32_t K, int32_t nacc, const float* A, const float* B, float* D,
                        int64_t* cycles);
/* stand-alone timing of one hot kernel on synthetic device data with an L2 flush between
 * iterations (bench.py roofline / profiles).  which: "softmax_ce" | "adam" | "embed_gather" |
 * "logits_gemm".  Returns the mean ms per launch and the algorithmic bytes / flops per launch. */
int  argsim_bench_kernel(argsim_handle*, const char* which, int64_t rows, int32_t iters,
                         float* ms, double* algo_bytes, double* algo_flops);
/* host-only (no GPU needed): the index pipeline of one batch -- trim (util_tf.py:40-57), length
 * sort, packed layout, lead/gold/mask construction (model.py:91-95) and the boolean_mask row
 * order (model.py:161,174).  Every output may be NULL.  keep: (b,T_tgt) uint8 or NULL (= keep all).
 * Outputs: len_src/len_tgt (b); counts[4] = {S, N, Tmax_src, Tmax_dec}; ids_src (S); lead, gold,
 * ref_row (N); enc_last (b); perm_src, perm_dec (b). Capacities are the caller's responsibility
 * (S <= b*T_src, N <= b*(T_tgt+1)). */
int  argsim_plan_batch(const int32_t* src, const int32_t* tgt, int32_t b, int32_t T_src, int32_t T_tgt,
                       int32_t bos, int32_t eos, const uint8_t* keep, int32_t* len_src, int32_t* len_tgt,
                       int64_t counts[4], int32_t* ids_src, int32_t* lead, int32_t* gold, int32_t* ref_row,
                       int32_t* enc_last, int32_t* perm_src, int32_t* perm_dec, char* err, int32_t err_cap);
/* host-only: src/model.py:75-80 evaluated in fp32 */
void argsim_schedule(int64_t step, float accelerate, float learn_rate, float* keepwd, float* anneal, float* update);

#ifdef __cplusplus
}
#endif
#endif
