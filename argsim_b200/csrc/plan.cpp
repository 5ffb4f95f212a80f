// plan.cpp -- see plan.h.  Host-only; compiled into libargsim_b200.so and exercised without a
// GPU through argsim_plan_batch (tests/test_plan.py compares it with the numpy oracle bit for bit).
#include "plan.h"
#include "philox.h"
#include <algorithm>
#include <math.h>
#include <numeric>

void SeqPlan::build(const std::vector<int>& steps_per_row) {
    steps = steps_per_row;
    b = (int)steps.size();
    perm.resize(b);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return steps[x] > steps[y]; });
    inv.assign(b, 0);
    for (int j = 0; j < b; ++j) inv[perm[j]] = j;
    Tmax = b ? steps[perm[0]] : 0;
    nact.assign(Tmax, 0);
    off.assign(Tmax + 1, 0);
    // nact[t] = #rows with steps > t; rows are sorted, so walk the sorted list once
    int j = b;
    for (int t = 0; t < Tmax; ++t) {
        while (j > 0 && steps[perm[j - 1]] <= t) --j;
        nact[t] = j;
    }
    for (int t = 0; t < Tmax; ++t) off[t + 1] = off[t] + nact[t];
    rows = off[Tmax];
}

// trim() of src/util_tf.py:54-57 on one batch-major row: counts ALL non-eos entries.  The
// reference assumes "any number of non-eos followed by any number of eos" (util_tf.py:50-52);
// we verify that contract instead of silently diverging from it.
static int trim_row(const int32_t* row, int T, int eos, bool* contract_ok) {
    int n = 0;
    for (int t = 0; t < T; ++t) n += (row[t] != eos);
    for (int t = 0; t < n; ++t)
        if (row[t] == eos) { *contract_ok = false; break; }
    return n;
}

std::string build_batch_plan(const int32_t* src, int b, int T_src, const int32_t* tgt, int T_tgt, int bos, int eos,
                             int need_dec, const DropoutSpec& drop, BatchPlan* P) {
    if (b <= 0) return "empty batch";
    if (!src || T_src <= 0) return "src is empty";
    P->b = b;
    bool ok = true;
    P->len_src.resize(b);
    for (int i = 0; i < b; ++i) {
        P->len_src[i] = trim_row(src + (size_t)i * T_src, T_src, eos, &ok);
        // model.py:135 gathers hs[len-1]; len 0 indexes -1 (undefined in the reference; data prep drops
        // empty posts, src/data_iac.py:28)
        if (P->len_src[i] < 1) return "src row " + std::to_string(i) + " has no non-eos token (len_src must be >= 1)";
    }
    if (!ok) return "src violates the trim() contract: eos inside a sequence (src/util_tf.py:50-52)";
    P->enc.build(P->len_src);
    const SeqPlan& E = P->enc;
    P->ids_src.resize(E.rows);
    for (int t = 0; t < E.Tmax; ++t)
        for (int j = 0; j < E.nact[t]; ++j) P->ids_src[E.off[t] + j] = src[(size_t)E.perm[j] * T_src + t];
    P->enc_last.resize(b);
    for (int i = 0; i < b; ++i) P->enc_last[i] = E.off[P->len_src[i] - 1] + E.inv[i];
    if (!need_dec) {
        P->dec = SeqPlan();
        P->ids_lead.clear(); P->labels.clear(); P->ref_row.clear(); P->len_tgt.clear();
        return "";
    }
    if (!tgt || T_tgt < 0) return "tgt is missing";
    P->len_tgt.resize(b);
    std::vector<int> dsteps(b);
    for (int i = 0; i < b; ++i) {
        P->len_tgt[i] = T_tgt ? trim_row(tgt + (size_t)i * T_tgt, T_tgt, eos, &ok) : 0;
        dsteps[i] = P->len_tgt[i] + 1;  // msk_tgt = [True] ++ not_eos  (model.py:91)
    }
    if (!ok) return "tgt violates the trim() contract: eos inside a sequence (src/util_tf.py:50-52)";
    P->dec.build(dsteps);
    const SeqPlan& D = P->dec;
    P->ids_lead.resize(D.rows);
    P->labels.resize(D.rows);
    for (int t = 0; t < D.Tmax; ++t) {
        for (int j = 0; j < D.nact[t]; ++j) {
            const int i = D.perm[j];
            const int32_t* row = tgt + (size_t)i * T_tgt;
            int lead;
            if (t == 0) {
                lead = bos;  // padded on AFTER the dropout multiply (model.py:94-95): never dropped
            } else {
                lead = row[t - 1];
                if (drop.train) {
                    int keep;
                    if (drop.keep) {
                        keep = drop.keep[(size_t)i * T_tgt + (t - 1)] != 0;
                    } else {
                        uint32_t c[4] = {(uint32_t)(drop.rows ? drop.rows[i] : drop.row0 + i), (uint32_t)(t - 1), (uint32_t)PHILOX_STREAM_KEEP,
                                         (uint32_t)(drop.seed >> 32)};
                        philox4x32_10(c, (uint32_t)drop.seed, (uint32_t)drop.step);
                        keep = u01_24(c[0]) < drop.rate_keepwd;  // tf.random_uniform(...) < rate_keepwd
                    }
                    lead *= keep;  // dropped -> 0 == unk
                }
            }
            P->ids_lead[D.off[t] + j] = lead;
            P->labels[D.off[t] + j] = (t < P->len_tgt[i]) ? row[t] : eos;  // gold = tgt ++ [eos]
        }
    }
    // reference row order of tf.boolean_mask(h, msk_tgt): time-major, original batch order
    P->ref_row.resize(D.rows);
    for (int t = 0; t < D.Tmax; ++t) {
        int r = D.off[t];
        for (int i = 0; i < b; ++i)
            if (dsteps[i] > t) P->ref_row[D.off[t] + D.inv[i]] = r++;
    }
    return "";
}

void schedule_f32(int64_t step, float accelerate, float learn_rate, float* keepwd, float* anneal, float* update) {
    const float rate = accelerate * (float)step;
    if (keepwd) *keepwd = 1.0f / (1.0f + expf(-rate));
    if (anneal) *anneal = tanhf(rate);
    if (update) *update = learn_rate / (sqrtf(rate) + 1.0f);
}
