// engine.h -- the sequence-VAE step (src/model.py:75-189) as a device program over packed,
// length-sorted rows.  Owns parameters, Adam slots, the activation arena, streams and the
// (optional) NCCL communicator.  The C ABI in capi.cu is a thin shell around this class.
#pragma once
#include "../../include/argsim_b200.h"
#include "kernels.h"
#include "plan.h"
#include <map>
#include <string>
#include <vector>

struct ParamInfo {
    std::string name;
    int rank;
    int64_t shape[2];
    size_t off;   // element offset in the flat buffers
    size_t n;
};

class Arena {
public:
    char* base = nullptr;
    size_t cap = 0, used = 0, high = 0;
    bool dry = false;
    void reset(bool dry_) { used = 0; dry = dry_; }
    void* alloc(size_t bytes) {
        size_t o = (used + 255) & ~size_t(255);
        used = o + bytes;
        if (used > high) high = used;
        if (dry) return (void*)(uintptr_t)(o + 256);  // never dereferenced
        if (used > cap) throw std::runtime_error("activation arena overflow (internal error)");
        return base + o;
    }
};

struct KTimer {
    std::string name;
    cudaEvent_t a, b;
    double gflop = 0.0;   // algorithmic work of the bracketed launches (GEMM groups), 0 if the caller accounts for it
};

class Engine {
public:
    explicit Engine(const argsim_config& c);
    ~Engine();

    argsim_config cfg;
    int V, D, R, L, H;
    bool is_bf16;    // precision mode
    bool use_tc;     // tcgen05 GEMMs
    bool use_mma;    // persistent GRU kernels
    std::string err;

    std::vector<ParamInfo> params;
    std::map<std::string, int> pindex;
    size_t nflat = 0;
    float *p = nullptr, *g = nullptr, *m = nullptr, *v = nullptr;
    bf16* ph = nullptr;
    int64_t step = 0;
    uint64_t seed = 0;

    void init_params(uint64_t seed);
    void set_param(const std::string& name, const float* src);
    void get_flat(const float* flat, const std::string& name, float* dst);
    void set_flat(float* flat, const std::string& name, const float* src);
    void refresh_shadow(size_t off, size_t n);
    void copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);

    // --- steps (host buffers in/out, blocking) ---
    void train_step(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, const uint8_t* keep, const float* eps,
                    int64_t n_tok_global, int64_t b_global, int64_t row0, bool apply_update, argsim_step_stats* out);
    void eval_step(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, float* errt, float* lgen, int64_t cap,
                   float* lkld, int64_t* n_rows, int32_t* pred);
    void embed(const int32_t* src, int b, int T, float* mu_out);
    void embed_one(const int32_t* src, int b, int T, float* mu_out);
    struct DecodeBufs;
    DecodeBufs* dbuf = nullptr;
    DecodeBufs& decode_bufs(int b, int steps);
    void decode_seed(DecodeBufs& B);
    void decode_state_changed(DecodeBufs& B);
    void decode_one(DecodeBufs& B);
    void decode_init(const float* z, int b, float* state);
    int decode_loop(const float* z, int b, int steps, int32_t* tokens);   // returns t <= steps; tokens is (b, steps) row-major
    void decode_step(const int32_t* lead, int b, float* state, int32_t* pred);
    // Pipelined form of train_step (src/train.py:118 fetches nothing but the op, so the call need not block): submit
    // enqueues the step and returns; wait returns the oldest un-waited step's statistics. At most two steps may be
    // un-waited, so the host plan + H2D of step n+1 overlap the device's step n. train_step == submit + wait.
    void train_step_submit(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, const uint8_t* keep, const float* eps,
                           int64_t n_tok_global, int64_t b_global, int64_t row0, bool apply_update);
    void train_step_wait(argsim_step_stats* out);
    // global row index of every row of the NEXT submitted train step (consumed by it): keys the Philox streams of word
    // dropout and eps, so that un-injected randomness does not depend on how the batch is dealt over the ranks
    void set_global_rows(const int64_t* rows, int b);
    std::vector<int64_t> next_rows;
    void drain();   // every submitted step complete on the device (their statistics stay queued for train_step_wait)
    void bench_resident(int iters, float* ms);
    void save(const char* path);
    void load(const char* path);

    bool nvtx = false;   // argsim_profiler(on): NVTX range per device program, NVTX mark at every phase boundary
    std::vector<std::string> tnames;
    std::vector<float> tms;
    long long launches0 = 0;

private:
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};  // main, second GRU direction, comm
    cudaStream_t sw[8] = {};                            // decoder wavefront: one stream per layer
    cudaStream_t sad = nullptr;                         // data parallel: the early Adam update (behind the layer-2 bucket's all-reduce)
    cudaStream_t swg = nullptr;                         // lowest-priority side stream: weight-gradient GEMMs and bias column sums
                                                        // run on the SMs the recurrence launches of the layer below leave free
    int group_cap = 0;                                  // ARGSIM_GROUP_CAP: groups of 16 CTAs the slice planners may use (0 = 9, or 8 under data parallel)
    int side_units = 32;                                // ARGSIM_SIDE_UNITS: k-blocks per work unit of the side stream's GEMMs (0 = persistent CTAs; 32: 10.50 -> 10.41 ms/step)
    int enc_bwd_chunk = 8;                              // 8-row chunks in the encoder's BPTT segment launches (ARGSIM_ENC_BWD_CHUNK=16: 16-row chunks there too).
                                                        // Encoder BPTT 3.53 -> 3.36 ms, step 10.34 -> 10.15 ms; the decoder wavefront is slower with them (1.33 -> 1.42 /
                                                        // 1.86 -> 1.92 ms) and keeps 16-row chunks
    int dec_early_on = 1;                               // ARGSIM_DEC_EARLY=0: decoder layer-0 gather + projection on the main stream after the encoder (10.66 vs 10.64 ms/step)
    bool slice_budget = true;                           // ARGSIM_NO_SLICE_BUDGET: every wavefront launch takes 16-row slices
    bool early_adam = true;                             // ARGSIM_NO_EARLY_ADAM: one Adam launch after the last gradient
    bool seg_wgrad_on = true;                           // ARGSIM_NO_SEG_WGRAD: last encoder layer's weight gradients in one piece
    int logit_chunk = 0;                                // ARGSIM_LOGIT_CHUNK: rows per vocabulary-projection chunk (0 = 4096 / 2048)
    int wgrad_overlap = 1;                              // ARGSIM_WGRAD_OVERLAP=0: weight gradients on the main stream, in line
    std::vector<cudaEvent_t> evpool;
    size_t evcount = 0;
    cudaEvent_t next_event();
    cudaEvent_t ev_embed[2] = {nullptr, nullptr};
    cudaEvent_t next_event_persistent(int i);
    int dec_seg = 64;                                   // time steps per wavefront segment (0 = off)
    int enc_seg = 171;                                  // encoder BPTT: time steps per (direction, segment) launch (0 = one launch per layer)
    int pad_wave = 0;                                   // 2: concurrent segment launches are padded to 128 blocks (plain launch), 0: not
    bool enc_seg_fwd = false;                           // measured: the forward pass is faster as one launch per layer (2.39 vs 2.56 ms), BPTT is not
    bool enc_segmented(const SeqPlan& E) const;
    void enc_slice_plan(const SeqPlan& E, int nseg, bool bptt, std::vector<int>* want8) const;
    bool dec_wavefront(const SeqPlan& Dp) const;
    cudaEvent_t ev_bucket = nullptr, ev_comm = nullptr;
    Arena arena;
    GruMmaCtx* mma = nullptr;
    GruTcCtx* tc = nullptr;      // tensor-memory recurrence (gru_tc.cu)
    int gru_tc_mode = 4;         // ARGSIM_GRU_TC: bit 0 = all forward, bit 1 = all backward recurrences on the tcgen05 kernels; bit 2 (default) =
                                 // whole-layer forward launches with many rows per slice (embedding batches, large training batches) take the
                                 // TMA-fed tcgen05 kernel, everything else the register-stationary mma.sync kernels
    void rec_bwd(const GruBwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact, cudaStream_t q, int t0,
                 int Tseg, int slot, int want8, int pad, int chunk);
    // one forward recurrence launch on whichever persistent kernel is selected (want8: the mma.sync kernel's slice request)
    void rec_fwd(const GruFwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact, cudaStream_t q, int t0,
                 int Tseg, int slot, int want8, int pad);
    void* nccl_comm = nullptr;

    // staged plan
    BatchPlan plan;
    int* h_stage = nullptr;  // pinned; two slots of stage_cap ints (one per step in flight)
    int* d_stage = nullptr;
    size_t stage_cap = 0;
    int stage_slot = -1;        // >= 0: staging slot to use instead of submit_seq & 1 (pipelined embedding micro-batches)
    struct PendingStep {
        cudaEvent_t done = nullptr;
        double* h_stats = nullptr;   // pinned, 4 doubles
        float keepwd = 0.f, anneal = 0.f, lr = 0.f;
        double n_glob = 0.0, b_glob = 0.0;
        int64_t step_after = 0;
    };
    PendingStep pend[2];
    uint64_t submit_seq = 0;   // steps submitted so far; slot = seq & 1
    int pend_n = 0;            // submitted and not yet waited for (oldest: slot (submit_seq - pend_n) & 1)
    struct DevPlan {
        int *ids_src, *ids_lead, *labels, *enc_last, *dec_perm, *enc_off, *enc_nact, *dec_off, *dec_nact, *row_ids;
    } dp{};
    float* d_eps_in = nullptr;  // injected eps staging (b,R)
    size_t eps_cap = 0;
    bool have_eps = false;
    double* d_stats = nullptr;  // 4 doubles
    double* h_stats = nullptr;  // pinned
    float* h_out = nullptr;     // pinned scratch for per-row outputs
    size_t h_out_cap = 0;

    // last staged step parameters (bench_resident replays them)
    struct StepArgs {
        int train = 0;
        int64_t n_tok_global = 0, b_global = 0, row0 = 0;
    } last;

    // phase timing
    std::vector<cudaEvent_t> pev;
    std::vector<std::string> pnames;
    size_t pcount = 0;
    void phase(const char* name);
    std::vector<KTimer> ktimers;
    size_t kcount = 0;
    void kbegin(const char* name, cudaStream_t q = nullptr);
    void kend(cudaStream_t q = nullptr, double gflop = 0.0);
    bool in_ktimer = false;
    bool decoding = false;      // inside decode_one: no per-GEMM timers (the launches may be captured into a graph)
    bool tied = true;   // logit_use_embed (src/model.py:164-168)
    bool dp_one_allreduce = false;   // ARGSIM_DP_ONE_ALLREDUCE: one all-reduce behind the backward pass instead of overlapped buckets
    bool attentive = false;   // src/model.py:136-145
    static constexpr int ATT_HEADS = 8;   // attention(..., head=8), src/model.py:18
    int enc_kind = 0;   // 0: stacked bidirectional (config.json); 1: two L-layer stacks, concatenated at the top; 2: one stack
    int EH = 0;         // encoder output width: 2H (bidirectional) or H
    std::string enc_prefix(int d, int j) const {   // encoder branches 1 / 2 (src/model.py:124-131)
        return std::string(enc_kind == 1 ? (d ? "encode/rnn/bwd/l" : "encode/rnn/fwd/l") : "encode/rnn/l") + std::to_string(j) + "/";
    }
    void collect_timings();

    Mat pmat(const std::string& name);
    const ParamInfo& pinfo(const std::string& name) const;
    float* gptr(const std::string& name) { return g + pinfo(name).off; }
    Mat gmat(const std::string& name);

    void stage(const int32_t* src, const int32_t* tgt, int b, int Ts, int Tt, int need_dec, const DropoutSpec& drop,
               const float* eps);
    // device program; mode: 0 = embed (mu only), 1 = valid, 2 = train(+backward)
    struct Out {
        float* mulv = nullptr;       // (b,2R)
        float* kld_samp = nullptr;   // (b,R)
        float* loss_samp = nullptr;  // (N)
        float* err_samp = nullptr;
        int* pred = nullptr;
    } outp;
    void run_device(int mode, bool apply_update);
    void ensure_arena(int mode);
    void program(int mode, bool apply_update);

    Mat act(long long rows, int cols);    // operand-typed activation (fp32 view or bf16 view by mode)
    Mat f32(long long rows, int cols);
    Mat both(long long rows, int cols);
    void gemm(const Mat& A, int a_mn, const Mat& B, int b_mn, const Mat& C, long long M, int N, long long K, float alpha,
              const float* bias, int accumulate, cudaStream_t q = nullptr);
    void colsum(const Mat& A, long long rows, int cols, float* out, int accumulate = 0, cudaStream_t q = nullptr);
    void gather_embed(const int* ids, long long n, const Mat& out, cudaStream_t q = nullptr);
    void gru_fwd(GruFwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact);
    void gru_bwd(GruBwdArgs* dirs, int ndir, const SeqPlan& P, const int* d_off, const int* d_nact);
    void allreduce_bucket(size_t off0, size_t off1, cudaStream_t after = nullptr);   // after: the stream whose work produced the bucket (default main)
    float* gru_work = nullptr;
    size_t gru_work_cap = 0;
    size_t bucket_lo = 0;  // grads below this offset are already reduced
};
