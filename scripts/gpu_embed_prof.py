import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from argsim_b200 import _lib
from argsim_b200.synth import synth_batch
CFG = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, flags=_lib.FLAG_KERNEL_TIMERS, **CFG)
h.init_params(0)
data = synth_batch(4096, 'ibm', 8192, seed=0)
lens = (data != 1).sum(1)
data = np.ascontiguousarray(data[:, :int(lens.max())])
print('tokens', int(lens.sum()), 'max len', int(lens.max()), 'mean', float(lens.mean()))
for _ in range(2):
    h.embed(data)
t0 = time.perf_counter(); mu = h.embed(data); t1 = time.perf_counter()
print('embed 4096 e2e ms', (t1 - t0) * 1e3)
order = np.argsort(-lens, kind='stable')
for i0 in range(0, 4096, 1024):
    sub = data[order[i0:i0 + 1024]]
    sub = np.ascontiguousarray(sub[:, :int((sub != 1).sum(1).max())])
    h.embed(sub)
    t0 = time.perf_counter(); h.embed(sub); t1 = time.perf_counter()
    tm = h.last_timings()
    print('micro-batch', i0, 'T', sub.shape[1], 'tokens', int((sub != 1).sum()), 'e2e ms %.2f' % ((t1 - t0) * 1e3), {k: round(v, 3) for k, v in tm.items() if '#' not in k})
