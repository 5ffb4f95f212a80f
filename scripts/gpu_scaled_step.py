"""two training steps at BASELINE configs[4] dimensions (D = 2048, V = 32768, R = 4096), 64 rows of at most 24 tokens: the
same per-launch shapes of the wide-model recurrence kernels as the full configs[4] step, short enough for ncu --set full"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from argsim_b200 import _lib
cfg = dict(dim_tgt=32768, dim_emb=2048, dim_rep=4096, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, **cfg)
h.init_params(0)
rng = np.random.default_rng(0)
src = rng.integers(3, 32768, (64, 25)).astype(np.int32)
src[:, -1] = 1
for _ in range(2):
    st = h.train_step(src, src)
print(st['loss'], h.launch_count())
