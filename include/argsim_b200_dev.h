/* argsim_b200_dev.h -- C ABI of libargsim_b200_dev.so: development microbenchmarks that are NOT part of the product
 * library (libargsim_b200.so, include/argsim_b200.h).  Nothing on the hot path links or loads this library; it exists so
 * that the measurements DESIGN.md section 5 quotes can be reproduced (scripts/gpu_xbench.py, scripts/gpu_xbench3.py). */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* cycles per all-gather round among the 16 CTAs of a recurrence group (every CTA publishes rows x 32 bf16 units and
 * needs all 512 before going on), `groups` groups running side by side.  method 0: L2 words with in-band tags, volatile;
 * 1: the same, relaxed.gpu; 2: cluster, tagged words into the peers' shared memory; 3: cluster, st.async + remote
 * mbarrier; 4: cluster, one cp.async.bulk per peer issued by 16 threads of 8 warps; 5: the same issued by one warp.
 * +16: groups formed by physical SM id, +32: grid padded to 8 groups, +64: six leading padding blocks (methods 0 / 1).
 * Returns 0, or -2 with the message in argsim_dev_last_error(). */
int  argsim_bench_exchange(int32_t device, int32_t method, int32_t groups, int32_t rows, int32_t iters,
                           double* cycles_per_round, int32_t* max_clusters_or_null);
const char* argsim_dev_last_error(void);

#ifdef __cplusplus
}
#endif
