import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from conftest import SMALL, ragged_batch
from oracle import vae_oracle as O
from test_gpu_parity import _mk, _inject, _oracle_keep
from argsim_b200 import _lib

def err(mu, ref):
    return float(np.sqrt(((mu - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))

cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
for variant in sys.argv[1:]:
    L = int(variant[1]) if variant[0] == 'L' else 3
    c = dict(cfg, rnn_layers=L)
    h, P = _mk(c, _lib.BF16, flags=4)
    for b, ts, seed in ((10, 20, 50), (2, 5, 1), (10, 20, 50), (64, 40, 2)):
        src = ragged_batch(b, ts, c['dim_tgt'], seed)
        ov, _ = O.forward(P, c, src, src, 'valid')
        print(variant, 'b', b, 'S', int((src != 1).sum()), 'embed err', err(h.embed(src), ov['mu']), flush=True)
        e = h.eval_step(src, src)
        print(variant, '   eval kld', float(e['loss_kld_samp'].mean()), float(ov['loss_kld']), 'gen', float(e['loss_gen_samp'].mean()), float(ov['loss_gen']), flush=True)
