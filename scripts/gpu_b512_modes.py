"""b = 512 training step (BASELINE configs[2] on one GPU): phase times under the recurrence kernel choices"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from argsim_b200 import _lib
from argsim_b200.synth import synth_batch
CFG = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, flags=_lib.FLAG_KERNEL_TIMERS, **CFG)
h.init_params(0)
for b in [int(x) for x in os.environ.get("BS", "128,256,512").split(",")]:
    src = synth_batch(b, 'iac', 8192, seed=0)
    for _ in range(3):
        st = h.train_step(src, src)
    tm = h.last_timings()
    ph = {k: round(v, 2) for k, v in tm.items() if not k.startswith('k:')}
    print('GRU_TC', os.environ.get('ARGSIM_GRU_TC', 'default'), 'b', b, 'total %.2f' % sum(ph.values()), ph, flush=True)
