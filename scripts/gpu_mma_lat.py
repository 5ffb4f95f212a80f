"""cycles of a K = 512 TS-form MMA chain (32 tcgen05.mma, M = 128) as a function of N and of the number of independent
accumulators the instructions rotate over (evidence for DESIGN section 5)."""
import json, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
from argsim_b200 import _lib
rng = np.random.default_rng(0)
out = {}
for N in (16, 32, 64, 128):
    A = rng.standard_normal((128, 512)).astype(np.float32)
    B = rng.standard_normal((N, 512)).astype(np.float32)
    for nacc in (1, 2, 4, 8, 16):
        if nacc * N > 256:
            continue
        D, cyc = _lib.test_ts_mma(A, B, nacc=nacc, want_cycles=True)
        out['N%d_acc%d' % (N, nacc)] = dict(total=cyc & 0xffffffff, issue=cyc >> 32)
        print('N', N, 'accumulators', nacc, 'cycles for 32 MMAs: issue -> mbarrier', cyc & 0xffffffff, ' issue loop', cyc >> 32)
for K in (64, 128, 256):
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((16, K)).astype(np.float32)
    D, cyc = _lib.test_ts_mma(A, B, nacc=1, want_cycles=True)
    out['N16_K%d_acc1_rolled' % K] = dict(total=cyc & 0xffffffff, issue=cyc >> 32)
    print('N 16 K', K, 'one accumulator, rolled loop:', cyc & 0xffffffff, cyc >> 32)
json.dump(out, open('gpurun_out/r2_ts_mma_latency.json', 'w'), indent=1)
