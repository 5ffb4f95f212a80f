#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 900 python - <<'PY' 2>&1 | tail -8
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.getcwd())
from argsim_b200 import _lib
cfg = dict(dim_tgt=32768, dim_emb=2048, dim_rep=4096, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, flags=_lib.FLAG_KERNEL_TIMERS, **cfg)
h.init_params(0); h.set_seed(0)
rng = np.random.default_rng(0)
b = 64
src = rng.integers(3, cfg['dim_tgt'], (b, 512)).astype(np.int32)
t0 = time.perf_counter(); st = h.train_step(src, src); t1 = time.perf_counter()
print('C5 first step', round(t1 - t0, 3), 's', st['loss'], st['loss_kld'])
t0 = time.perf_counter(); st = h.train_step(src, src); t1 = time.perf_counter()
print('C5 second step (e2e)', round(t1 - t0, 3), 's', st['loss'])
ms = h.bench_resident(2)
tm = h.last_timings()
print('C5 resident ms/step', round(ms, 2), 'seq/s', round(b / ms * 1e3, 1))
print({k: round(v, 2) for k, v in tm.items() if not k.endswith('#n') and not k.endswith('#gflop')})
json.dump(dict(ms_per_step=ms, seq_per_s=b / ms * 1e3, timings=tm, stats=st), open('gpurun_out/c5_scaled.json', 'w'))
PY
