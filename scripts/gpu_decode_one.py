import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from argsim_b200 import _lib
cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 512
z = np.random.default_rng(0).standard_normal((b, 1024)).astype(np.float32)
h = _lib.Handle(precision=_lib.BF16, **cfg)
h.init_params(0)
tok = h.decode(z, steps=3)
print(tok.shape)
