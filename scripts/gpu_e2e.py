import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from argsim_b200 import _lib
from argsim_b200.synth import synth_batch
cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, **cfg)
h.init_params(0); h.set_seed(0)
src = synth_batch(64, 'iac', 8192, seed=0)
rng = np.random.default_rng(0)
keep = (rng.random(src.shape) < 0.5).astype(np.uint8)
eps = rng.standard_normal((64, 1024)).astype(np.float32)
def run(n, **kw):
    for _ in range(3): h.train_step(src, src, **kw)
    t0 = time.perf_counter()
    for _ in range(n): h.train_step(src, src, **kw)
    return (time.perf_counter() - t0) / n * 1e3
print('e2e ms/step, philox on host      :', round(run(30), 3))
print('e2e ms/step, injected keep + eps :', round(run(30, keep=keep, eps=eps), 3))
print('resident ms/step                 :', round(h.bench_resident(30), 3))
t0 = time.perf_counter()
for _ in range(100): _lib.plan_batch(src, src)
print('plan_batch (no dropout) ms       :', round((time.perf_counter() - t0) / 100 * 1e3, 3))
