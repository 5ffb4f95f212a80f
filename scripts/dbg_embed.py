import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from conftest import SMALL, ragged_batch
from oracle import vae_oracle as O
from test_gpu_parity import _mk
from argsim_b200 import _lib

def err(mu, ref):
    return float(np.sqrt(((mu - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))

for name, cfg in (('small', dict(SMALL)), ('L1', dict(SMALL, rnn_layers=1)), ('L2', dict(SMALL, rnn_layers=2))):
    h, P = _mk(cfg, _lib.BF16, flags=4)
    hf, _ = _mk(cfg, _lib.FP32_VALIDATE)
    for b, tmax, seed, tmin in ((1, 5, 0, 5), (2, 5, 0, 5), (4, 8, 0, 8), (9, 14, 30, 1), (9, 14, 50, 1), (9, 14, 50, 14), (16, 14, 3, 1), (40, 14, 3, 1), (9, 3, 5, 1)):
        src = ragged_batch(b, tmax, cfg['dim_tgt'], seed, tmin=tmin)
        ov, _ = O.forward(P, cfg, src, src, 'valid')
        print(name, 'b', b, 'tmax', tmax, 'seed', seed, 'S', int((src != 1).sum()), 'bf16 err %.4f' % err(h.embed(src), ov['mu']),
              'fp32 err %.2e' % err(hf.embed(src), ov['mu']), 'lens', sorted((src != 1).sum(1).tolist(), reverse=True)[:12])
