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

cfg = dict(SMALL)
src = ragged_batch(9, 14, cfg['dim_tgt'], 50)
tgt = ragged_batch(9, 12, cfg['dim_tgt'], 51)
keep, eps = _inject(cfg, tgt, 52)
for label, step, first_embed, prec in (('A', 20000, False, _lib.BF16), ('B', 0, False, _lib.BF16), ('C', 20000, True, _lib.BF16), ('D', 20000, False, _lib.FP32_VALIDATE),
                                       ('E', 3000, False, _lib.BF16)):
    h, P = _mk(cfg, prec, flags=4)
    ov, _ = O.forward(P, cfg, src, tgt, 'valid')
    h.step = step
    if first_embed:
        print(label, 'embed first', err(h.embed(src), ov['mu']))
    st = h.grad_step(src, tgt, keep=keep, eps=eps)
    print(label, 'step', step, 'kld', st['loss_kld'], float(ov['loss_kld']), 'gen', st['loss_gen'])
    print(label, 'embed after grad', err(h.embed(src), ov['mu']))
    p1 = h.get_params()
    print(label, 'params changed:', [k for k in P if not np.array_equal(P[k].astype(np.float32), p1[k])][:5])
    h.set_params({k: v.astype(np.float32) for k, v in P.items()})
    print(label, 'embed after reset params', err(h.embed(src), ov['mu']))
    st = h.grad_step(src, tgt, keep=keep, eps=eps)
    print(label, 'kld again', st['loss_kld'])
