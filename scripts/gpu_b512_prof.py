"""b = 512: in-kernel phase cycles of the tcgen05 BPTT kernel (ARGSIM_GRU_TC=3 ARGSIM_GRU_PROF=1)"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from argsim_b200 import _lib
from argsim_b200.synth import synth_batch
CFG = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, **CFG)
h.init_params(0)
src = synth_batch(512, 'iac', 8192, seed=0)
h.train_step(src, src)
sys.stderr.write('=== second step\n')
h.train_step(src, src)
