import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from argsim_b200 import _lib
cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
for b in (128, 192, 256, 320, 512):
    z = np.random.default_rng(0).standard_normal((b, 1024)).astype(np.float32)
    h = _lib.Handle(precision=_lib.BF16, **cfg)
    h.init_params(0)
    h.decode(z, steps=8)
    t0 = time.perf_counter(); tok = h.decode(z, steps=64); dt = time.perf_counter() - t0
    print(os.environ.get('ARGSIM_STEP_NO_PDL', '-'), 'b', b, 'steps', tok.shape[1], 'ms/step %.3f' % (dt / max(tok.shape[1], 1) * 1e3), flush=True)
    h.close()
