"""Independent torch-CPU build of the same graph (src/model.py:75-189) with autograd.

TEST INFRASTRUCTURE / REPORTED CPU BASELINE ONLY (see oracle/vae_oracle.py header;
parity unpinned against TF).  Two uses:
  * cross-check of the numpy oracle's forward and analytic backward (tests/test_oracle.py);
  * the "port" CPU baseline that bench.py times on the GPU box's host cores
    (torch.gru = the library GRU, MKL/oneDNN GEMMs, all host threads), i.e. the
    closest runnable stand-in for the reference's TF path, which cannot run here.
It runs the GRUs over all padded steps exactly as the reference does (docs/log.org:107-112).
"""
import numpy as np
import torch


def _gru(x, h0, W, R, bW, bR):
    """1-layer time-major GRU via the torch library kernel (gate order r,z,n == cuDNN r,u,n)."""
    out, hT = torch._VF.gru(x, h0[None], [W, R, bW, bR], True, 1, 0.0, False, False, False)
    return out, hT[0]


def reverse_sequence(x, lens):
    T = x.shape[0]
    t = torch.arange(T)[:, None]
    idx = torch.where(t < lens[None, :], lens[None, :] - 1 - t, t)
    return x.gather(0, idx[..., None].expand_as(x))


def to_torch(P, dtype=torch.float32, requires_grad=False):
    return {k: torch.tensor(np.asarray(v), dtype=dtype).requires_grad_(requires_grad) for k, v in P.items()}


def forward(P, cfg, src, tgt, mode='valid', step=0, keep=None, eps=None, encoder_only=False):
    bos, eos = cfg.get('bos', 2), cfg.get('eos', 1)
    L, D = cfg.get('rnn_layers', 3), cfg.get('dim_emb', 512)
    H = D
    E = P['embed/embedding']
    dt = E.dtype
    rate = torch.tensor(cfg.get('accelerate', 1e-4) * float(step), dtype=dt)
    anneal = torch.tanh(rate)
    lr = cfg.get('learn_rate', 1e-3) / (torch.sqrt(rate) + 1.0)

    src_tm = torch.as_tensor(np.asarray(src, np.int64)).t()
    ne = src_tm != eos
    len_src = ne.sum(0)
    src_tm = src_tm[:int(len_src.max())]
    b = src_tm.shape[1]
    x = E[src_tm]
    z0 = torch.zeros(b, H, dtype=dt)
    bidir, stacked = cfg.get('bidirectional', True), cfg.get('bidir_stacked', True)
    if bidir and stacked:                                   # model.py:118-122
        for i in range(1, L + 1):
            pf, pb = 'encode/rnn%d/fwd/' % i, 'encode/rnn%d/bwd/' % i
            fwd, _ = _gru(x, z0, P[pf + 'W'], P[pf + 'R'], P[pf + 'bW'], P[pf + 'bR'])
            bwd, _ = _gru(reverse_sequence(x, len_src), z0, P[pb + 'W'], P[pb + 'R'], P[pb + 'bW'], P[pb + 'bR'])
            x = torch.cat([fwd, reverse_sequence(bwd, len_src)], -1)
    else:                                                   # model.py:124-131
        outs = []
        for k, st in enumerate(['encode/rnn/fwd/', 'encode/rnn/bwd/'] if bidir else ['encode/rnn/']):
            y = reverse_sequence(x, len_src) if k == 1 else x
            for j in range(L):
                p = '%sl%d/' % (st, j)
                y, _ = _gru(y, z0, P[p + 'W'], P[p + 'R'], P[p + 'bW'], P[p + 'bR'])
            outs.append(reverse_sequence(y, len_src) if k == 1 else y)
        x = torch.cat(outs, -1)
    h = x[len_src - 1, torch.arange(b)]
    if cfg.get('attentive', False):                         # model.py:136-145, repaired as oracle/vae_oracle.py:cata_forward states
        s_, d_ = x.shape[0], x.shape[2]
        hd, c = 8, d_ // 8
        q = (h @ P['encode/cata/q/kernel'] + P['encode/cata/q/bias']).reshape(b, hd, c)
        k = (x @ P['encode/cata/k/kernel'] + P['encode/cata/k/bias']).reshape(s_, b, hd, c)
        v = (x @ P['encode/cata/v/kernel'] + P['encode/cata/v/bias']).reshape(s_, b, hd, c)
        a = torch.einsum('bhc,sbhc->bhs', q, k) * (c ** -0.5)
        pad = torch.arange(s_)[None, :] >= len_src[:, None]
        a = torch.softmax(a.masked_fill(pad[:, None, :], float('-inf')), -1)
        y = torch.einsum('bhs,sbhc->bhc', a, v).reshape(b, d_)
        xx = h + y @ P['encode/cata/p/kernel'] + P['encode/cata/p/bias']
        h = torch.nn.functional.layer_norm(xx, (d_,), P['encode/cata/LayerNorm/gamma'], P['encode/cata/LayerNorm/beta'], 1e-12)
    mu = h @ P['latent/mu/kernel'] + P['latent/mu/bias']
    o = dict(mu=mu, z=mu, rate_update=float(lr), rate_anneal=float(anneal))
    if encoder_only:
        return o
    lv = h @ P['latent/lv/kernel'] + P['latent/lv/bias']
    z = mu
    if mode == 'train':
        z = mu + torch.exp(0.5 * lv) * torch.as_tensor(np.asarray(eps), dtype=dt)
    hx = z @ P['latent/ex/kernel'] + P['latent/ex/bias']

    tgt_tm = torch.as_tensor(np.asarray(tgt, np.int64)).t()
    ne = tgt_tm != eos
    len_tgt = ne.sum(0)
    tmax = int(len_tgt.max())
    tgt_tm, ne = tgt_tm[:tmax], ne[:tmax]
    msk = torch.cat([torch.ones(1, b, dtype=torch.bool), ne], 0)
    gold = torch.cat([tgt_tm, torch.full((1, b), eos, dtype=torch.int64)], 0)
    lead = tgt_tm
    if mode == 'train':
        lead = lead * torch.as_tensor(np.asarray(keep, np.int64))
    lead = torch.cat([torch.full((1, b), bos, dtype=torch.int64), lead], 0)
    y = E[lead]
    flat = []
    for j in range(L):
        p = 'decode/rnn/l%d/' % j
        flat += [P[p + 'W'], P[p + 'R'], P[p + 'bW'], P[p + 'bR']]
    y, _ = torch._VF.gru(y, torch.stack((hx,) * L), flat, True, L, 0.0, False, False, False)
    hd = y[msk]
    ho = hd @ P['decode/out/kernel'] + P['decode/out/bias']
    if cfg.get('logit_use_embed', True):
        logits = ho @ ((D ** -0.5) * E.t())
    else:
        logits = ho @ P['logits/dense/kernel'] + P['logits/dense/bias']
    labels = gold[msk]
    gen = torch.nn.functional.cross_entropy(logits, labels, reduction='none')
    kld = 0.5 * (mu * mu + torch.exp(lv) - lv - 1.0)
    loss = anneal * kld.mean() + gen.mean()
    o.update(lv=lv, z=z, logits=logits, labels=labels, loss_gen_samp=gen, loss_kld_samp=kld,
             loss_gen=gen.mean(), loss_kld=kld.mean(), loss=loss,
             errt_samp=(logits.argmax(-1) != labels).float())
    return o


def adam_tf_(P, M, V, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    lr_t = lr * (1 - beta2 ** t) ** 0.5 / (1 - beta1 ** t)
    with torch.no_grad():
        for k, p in P.items():
            g = p.grad
            M[k].mul_(beta1).add_(g, alpha=1 - beta1)
            V[k].mul_(beta2).addcmul_(g, g, value=1 - beta2)
            p.addcdiv_(M[k], V[k].sqrt().add_(eps), value=-lr_t)
            p.grad = None


def train_step(P, M, V, cfg, src, tgt, step, keep, eps):
    """forward + backward + TF-form Adam; P tensors must have requires_grad."""
    o = forward(P, cfg, src, tgt, 'train', step, keep, eps)
    o['loss'].backward()
    adam_tf_(P, M, V, step + 1, o['rate_update'])
    return o
