import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from oracle import vae_oracle as O
from argsim_b200 import _lib
G = np.load(os.path.join(R, 'tests/golden/small_vae.npz'))
CFG = dict(dim_tgt=128, dim_emb=64, dim_rep=64, rnn_layers=2, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
P = {k[2:]: G[k] for k in G.files if k.startswith('P/')}
z = G['z_dec']
for trial in range(3):
    h = _lib.Handle(precision=_lib.FP32_VALIDATE, **CFG)
    h.set_params(P)
    s = h.decode_init(z)
    hx = z @ P['latent/ex/kernel'] + P['latent/ex/bias']
    print('trial', trial, 'init err', np.abs(s - hx[None]).max())
    st = [hx.copy(), hx.copy()]
    x = np.full(3, 2, np.int32)
    for step in range(3):
        xd, s = h.decode_step(x, s)
        y = P['embed/embedding'][x][None]
        for j in range(2):
            pre = 'decode/rnn/l%d/' % j
            y, _ = O.gru_forward(y, st[j], P[pre + 'W'], P[pre + 'R'], P[pre + 'bW'], P[pre + 'bR'])
            st[j] = y[0]
        lg = (y[0] @ P['decode/out/kernel'] + P['decode/out/bias']) @ (64 ** -0.5 * P['embed/embedding'].T)
        xo = lg.argmax(-1).astype(np.int32)
        print('  step', step, 'dev', xd, 'ora', xo, 'state err', [float(np.abs(s[j] - st[j]).max()) for j in range(2)],
              'top2 gap', np.sort(lg, -1)[:, -1] - np.sort(lg, -1)[:, -2])
        x = xo
