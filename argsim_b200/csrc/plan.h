// plan.h -- host-side index pipeline of one batch (pure C++, no CUDA): trim, length sort,
// packed time-major row layout, decoder lead/gold construction with word dropout, and the
// map back to the reference's boolean_mask row order.  Bit-exact integer work
// (reference: src/util_tf.py:40-57, src/model.py:84-95,135,161,174).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

// One set of b sequences laid out "packed": rows are sorted by step count (descending, stable)
// and stored time-major, so step t owns the contiguous rows [off[t], off[t] + nact[t]) and row
// off[t] + j belongs to sorted sequence j.  Padding never exists in HBM.
struct SeqPlan {
    int b = 0;
    int Tmax = 0;                // number of steps of the longest sequence
    std::vector<int> steps;      // (b) per ORIGINAL row
    std::vector<int> perm;       // (b) sorted position -> original row
    std::vector<int> inv;        // (b) original row -> sorted position
    std::vector<int> nact;       // (Tmax) rows active at step t
    std::vector<int> off;        // (Tmax+1) prefix sums of nact
    long long rows = 0;          // off[Tmax]
    void build(const std::vector<int>& steps_per_row);
};

struct BatchPlan {
    int b = 0;
    SeqPlan enc, dec;                   // encoder steps = len_src ; decoder steps = len_tgt + 1
    std::vector<int> len_src, len_tgt;  // (b) trim() lengths, original order
    std::vector<int> ids_src;           // (S) packed encoder token ids
    std::vector<int> ids_lead;          // (N) packed decoder inputs  [bos] ++ dropout(tgt)
    std::vector<int> labels;            // (N) packed decoder targets tgt ++ [eos]
    std::vector<int> enc_last;          // (b) packed encoder row of step len_src-1, per ORIGINAL row
    std::vector<int> ref_row;           // (N) packed decoder row -> row index in the reference's
                                        //     tf.boolean_mask order (time-major over original rows)
};

// Philox keep-mask parameters (used when no mask is injected).
struct DropoutSpec {
    int train = 0;               // 0: no dropout (valid / infer mode)
    const uint8_t* keep = nullptr;  // (b, T_tgt) injected mask or null
    float rate_keepwd = 1.f;
    uint64_t seed = 0, step = 0;
    long long row0 = 0;          // global index of row 0 (data parallel)
    const int64_t* rows = nullptr;  // (b) global index of every row (overrides row0 + i): the streams are then invariant
                                    // to how the global batch is dealt over the ranks
};

// Builds the plan; returns "" on success or an error message.  need_dec = 0 builds the encoder
// part only (embedding inference).
std::string build_batch_plan(const int32_t* src, int b, int T_src, const int32_t* tgt, int T_tgt, int bos, int eos,
                             int need_dec, const DropoutSpec& drop, BatchPlan* out);

// src/model.py:75-80 in fp32, exactly as TF evaluates it.
void schedule_f32(int64_t step, float accelerate, float learn_rate, float* keepwd, float* anneal, float* update);
