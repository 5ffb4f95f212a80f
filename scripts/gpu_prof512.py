"""per-phase cycles of the tensor-memory forward kernel in the throughput regime: b = 512 training step and b = 4096 embed"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from argsim_b200 import _lib
from argsim_b200.synth import synth_batch
CFG = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, flags=_lib.FLAG_KERNEL_TIMERS, **CFG)
h.init_params(0)
src = synth_batch(512, 'iac', 8192, seed=0)
for _ in range(2):
    st = h.train_step(src, src)
sys.stderr.write('=== b512 train step\n')
st = h.train_step(src, src)
tm = h.last_timings()
print('b512 phases', {k: round(v, 2) for k, v in tm.items() if not k.startswith('k:')})
data = synth_batch(4096, 'ibm', 8192, seed=0)
data = np.ascontiguousarray(data[:, :int((data != 1).sum(1).max())])
h.embed(data)
sys.stderr.write('=== embed 4096\n')
t0 = time.perf_counter(); h.embed(data); print('embed ms', (time.perf_counter() - t0) * 1e3)
