import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from conftest import SMALL, ragged_batch
from oracle import vae_oracle as O
from test_gpu_parity import _mk, _inject, _oracle_keep, rel
from argsim_b200 import _lib

def run(cfg, b, ts, tt, flags, step=20000):
    h, P = _mk(cfg, _lib.BF16, flags=flags)
    src = ragged_batch(b, ts, cfg['dim_tgt'], 50)
    tgt = ragged_batch(b, tt, cfg['dim_tgt'], 51)
    keep, eps = _inject(cfg, tgt, 52)
    h.step = step
    o, cache = O.forward(P, cfg, src, tgt, 'train', step=step, keep=_oracle_keep(keep, tgt, cfg['eos']), eps=eps.astype(np.float64))
    G = O.backward(P, cfg, cache)
    st = h.grad_step(src, tgt, keep=keep, eps=eps)
    print('flags', flags, {k: (st[k], float(o[k]), rel(st[k], o[k])) for k in ('loss', 'loss_gen', 'loss_kld')})
    for k in P:
        g = h.get_grad(k).astype(np.float64).ravel(); r = G[k].ravel()
        cos = g @ r / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
        print('  %-24s cos %.4f ratio %.4f' % (k, cos, np.linalg.norm(g) / (np.linalg.norm(r) + 1e-30)))
    mu = h.embed(src)
    ov, _ = O.forward(P, cfg, src, tgt, 'valid', step=step)
    print('  mu relerr max', np.abs(mu - ov['mu']).max() / np.abs(ov['mu']).max(), 'rms', np.sqrt(((mu-ov['mu'])**2).mean())/np.sqrt((ov['mu']**2).mean()))

cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
run(cfg, 10, 20, 17, 4)
run(dict(SMALL), 9, 14, 12, 4)
