"""CPU oracle for the argsim sequence-VAE hot path (numpy, fp64 truth / fp32 mode).

TEST INFRASTRUCTURE ONLY.  Nothing under ``argsim_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs use it, and only as the checker / reported baseline.

PARITY UNPINNED against the reference's own TF run: TensorFlow 1.x is not installed
here (nor on the GPU box) and the graph uses the GPU-only ``CudnnGRU`` op
(/root/reference/src/model.py:15), so the reference cannot be executed.  This file
restates the reference graph literally (padded, time-major, unsorted -- the way
src/model.py runs it) and is pinned against what the reference does publish:
  * the schedule table  docs/log.org:21-28            (tests/test_oracle.py)
  * the init bound      docs/log.org:86-88, src/model.py:109
  * the docstring shapes of trim / vpack / partition  src/util_tf.py:40-52, src/util_np.py:5-24
and cross-checked against an independent torch-CPU autograd build of the same
graph (oracle/vae_torch.py) and finite differences.

Library semantics restated (source not under /root/reference; public TF-1.x / cuDNN docs):
  * CudnnGRU cell  (model.py:15,118-122,160): gate order (r,u,n), two bias vectors,
        r = sig(W_r x + b_Wr + R_r h + b_Rr);  u = sig(W_u x + b_Wu + R_u h + b_Ru)
        n = tanh(W_n x + b_Wn + r*(R_n h + b_Rn));  h' = (1-u)*n + u*h
  * tf.layers.dense (model.py:12): y = x @ kernel + bias, kernel is (in, out)
  * sparse_softmax_cross_entropy_with_logits (model.py:180): logsumexp(l) - l[label]
  * tf.train.AdamOptimizer (model.py:189): beta1=.9 beta2=.999 eps=1e-8,
        lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)

Canonical parameter names / shapes (shared with the C ABI, include/argsim_b200.h):
  embed/embedding (V,D)
  encode/rnn{i}/{fwd,bwd}/{W (3H,in), R (3H,H), bW (3H), bR (3H)}   i = 1..L
  latent/{mu,lv}/{kernel (2H,R), bias (R)}   latent/ex/{kernel (R,D), bias (D)}
  decode/rnn/l{j}/{W (3H,D), R (3H,H), bW, bR}                      j = 0..L-1
  decode/out/{kernel (D,D), bias (D)}
"""
import numpy as np

# --------------------------------------------------------------------------------------
# index pipeline (bit exact integer work)
# --------------------------------------------------------------------------------------


def vpack(arrays, shape, fill, dtype=None):
    """src/util_np.py:5-13 -- rows of different length padded with `fill` at the end."""
    out = np.full(shape, fill, dtype)
    for i, a in enumerate(arrays):
        if i >= shape[0]:
            break
        out[i, :len(a)] = a
    return out


def partition(n, m, discard=False):
    """src/util_np.py:16-24 -- index pairs cutting range(n) into pieces of m."""
    pairs = [(i, i + m) for i in range(0, n - m + 1, m)]
    if n % m and not discard:
        pairs.append((n - n % m, n))
    return pairs


def sample(n, seed=0):
    """src/util_np.py:27-33 -- infinite index stream; the SAME list is reshuffled in
    place with the SAME seed each epoch (epoch k = k-fold composition of one permutation)."""
    data = list(range(n))
    while True:
        np.random.seed(seed)
        np.random.shuffle(data)
        yield from data


def trim(x_tm, eos):
    """src/util_tf.py:40-57 on a time-major int32 (T,b) array."""
    not_eos = x_tm != eos
    len_seq = not_eos.astype(np.int32).sum(axis=0).astype(np.int32)
    max_len = int(len_seq.max()) if len_seq.size else 0
    return x_tm[:max_len], not_eos[:max_len], len_seq


def decoder_io(tgt_tm, not_eos, bos, eos, keep=None):
    """src/model.py:91-95.  keep: bool/0-1 array (t-1,b) or None (valid/infer mode)."""
    b = tgt_tm.shape[1]
    msk_tgt = np.concatenate([np.ones((1, b), bool), not_eos], 0)
    gold = np.concatenate([tgt_tm, np.full((1, b), eos, np.int32)], 0)
    lead = tgt_tm.copy()
    if keep is not None:
        lead = lead * np.asarray(keep).astype(np.int32)
    lead = np.concatenate([np.full((1, b), bos, np.int32), lead], 0).astype(np.int32)
    return lead, gold.astype(np.int32), msk_tgt


def schedule(step, accelerate=1e-4, learn_rate=1e-3, dtype=np.float32):
    """src/model.py:75-80; step = global step BEFORE the update."""
    dt = np.dtype(dtype).type
    rate = dt(accelerate) * dt(step)
    keepwd = dt(1) / (dt(1) + np.exp(-rate))
    anneal = np.tanh(rate)
    update = dt(learn_rate) / (np.sqrt(rate) + dt(1))
    return dict(rate=rate, rate_keepwd=dt(keepwd), rate_anneal=dt(anneal), rate_update=dt(update))


def reverse_sequence(x, lens):
    """tf.reverse_sequence(seq_axis=0, batch_axis=1): reverse the first lens[b] steps of column b."""
    y = x.copy()
    for b, n in enumerate(lens):
        n = int(n)
        y[:n, b] = x[:n, b][::-1]
    return y


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------

def enc_stacks(cfg):
    """name prefixes of the encoder's independent GRU stacks for the non-default branches (src/model.py:124-131):
    non-stacked bidirectional = two L-layer stacks ('fwd' on the sequence, 'bwd' on the reversed one), unidirectional =
    one.  Canonical names: encode/rnn/{fwd,bwd}/l{j}/... and encode/rnn/l{j}/..."""
    if cfg.get('bidirectional', True) and cfg.get('bidir_stacked', True):
        return None
    return ['encode/rnn/fwd/', 'encode/rnn/bwd/'] if cfg.get('bidirectional', True) else ['encode/rnn/']


def param_shapes(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, logit_use_embed=True, bidirectional=True,
                 bidir_stacked=True, attentive=False, **_):
    V, D, R, L = dim_tgt, dim_emb, dim_rep, rnn_layers
    H = D
    shp = {'embed/embedding': (V, D)}
    stacks = enc_stacks(dict(bidirectional=bidirectional, bidir_stacked=bidir_stacked))
    if stacks is None:
        for i in range(1, L + 1):
            cin = D if i == 1 else 2 * H
            for d in ('fwd', 'bwd'):
                p = 'encode/rnn%d/%s/' % (i, d)
                shp[p + 'W'] = (3 * H, cin)
                shp[p + 'R'] = (3 * H, H)
                shp[p + 'bW'] = (3 * H,)
                shp[p + 'bR'] = (3 * H,)
    else:
        for st in stacks:
            for j in range(L):
                p = '%sl%d/' % (st, j)
                shp[p + 'W'] = (3 * H, D if j == 0 else H)
                shp[p + 'R'] = (3 * H, H)
                shp[p + 'bW'] = (3 * H,)
                shp[p + 'bR'] = (3 * H,)
    EH = 2 * H if bidirectional else H
    if attentive:   # src/model.py:33-35,45 (layer_aff v,k,q,p under encode/cata) and :139 (layer_nrm)
        for nm in 'qkvp':
            shp['encode/cata/%s/kernel' % nm] = (EH, EH)
            shp['encode/cata/%s/bias' % nm] = (EH,)
        shp['encode/cata/LayerNorm/gamma'] = (EH,)
        shp['encode/cata/LayerNorm/beta'] = (EH,)
    for nm, (i, o) in (('mu', (EH, R)), ('lv', (EH, R)), ('ex', (R, D))):
        shp['latent/%s/kernel' % nm] = (i, o)
        shp['latent/%s/bias' % nm] = (o,)
    for j in range(L):
        p = 'decode/rnn/l%d/' % j
        shp[p + 'W'] = (3 * H, D)
        shp[p + 'R'] = (3 * H, H)
        shp[p + 'bW'] = (3 * H,)
        shp[p + 'bR'] = (3 * H,)
    shp['decode/out/kernel'] = (D, D)
    shp['decode/out/bias'] = (D,)
    if not logit_use_embed:   # src/model.py:167-168: layer_aff(h, dim_tgt) = tf.layers.dense under scope 'logits'
        shp['logits/dense/kernel'] = (D, V)
        shp['logits/dense/bias'] = (V,)
    return shp


def init_params(cfg, seed=0, dtype=np.float64, bias_scale=0.0):
    """src/model.py:8-15,109-110 (A22): embedding U(+-sqrt(6/(V/D+1))); kernels glorot-uniform
    per canonical sub-matrix (per gate); biases zero (bias_scale>0 gives random biases so that
    parity tests exercise the bias paths)."""
    rng = np.random.default_rng(seed)
    V, D = cfg.get('dim_tgt', 8192), cfg.get('dim_emb', 512)
    H = D
    out = {}
    for name, shp in param_shapes(**cfg).items():
        if name == 'embed/embedding':
            bnd = (6.0 / (V / D + 1)) ** 0.5
            w = rng.uniform(-bnd, bnd, shp)
        elif name.endswith('LayerNorm/gamma'):   # tf.contrib.layers.layer_norm: scale starts at one
            w = 1.0 + (rng.uniform(-bias_scale, bias_scale, shp) if bias_scale else np.zeros(shp))
        elif len(shp) == 1:
            w = rng.uniform(-bias_scale, bias_scale, shp) if bias_scale else np.zeros(shp)
        elif name.endswith('/W') or name.endswith('/R'):
            cin = shp[1]
            bnd = (6.0 / (cin + H)) ** 0.5
            w = rng.uniform(-bnd, bnd, shp)
        else:
            bnd = (6.0 / (shp[0] + shp[1])) ** 0.5
            w = rng.uniform(-bnd, bnd, shp)
        out[name] = w.astype(dtype)
    return out


# --------------------------------------------------------------------------------------
# GRU (cuDNN form), padded time-major, exactly as the reference runs it
# --------------------------------------------------------------------------------------

def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


# --------------------------------------------------------------------------------------
# attentive=true (src/model.py:18-45,136-145)
# --------------------------------------------------------------------------------------
ATT_HEADS = 8        # attention(..., head=8), src/model.py:18
LN_EPS = 1e-12       # tf.contrib.layers.layer_norm -> tf.nn.batch_normalization(variance_epsilon=1e-12)


def cata_forward(P, h, hs, len_src):
    """src/model.py:136-145 as its docstring (:19-26) and comments state it, with the two defects that keep the reference
    from running repaired -- NOT a restatement of executable reference code (the branch is marked 'todo fixme' and
    config.json never enables it):
      * :33-35 apply tf.layers.dense to (b, d, s) tensors, i.e. over the TIME axis; the affines here act on the feature
        axis d, which is what the shape comments 'bds <- bvs' / 'bdt <- bqt' ask for;
      * :35 reshapes q to (b,h,c,s) although it has t (=1) positions; here (b,h,c,t).
    query = the final state h (b,d), t = 1; keys / values = all outputs hs (s,b,d); mask = log(msk_src) = 0 on the
    sequence's own steps, -inf on padding; 8 heads of c = d/8, scores scaled by c^-1/2; h <- layer_norm(h + p(y)) with
    the layer norm over the feature axis (biased variance, epsilon 1e-12, gamma and beta).
    Returns (h_out, cache)."""
    s, b, d = hs.shape
    hd, c = ATT_HEADS, d // ATT_HEADS
    q = h @ P['encode/cata/q/kernel'] + P['encode/cata/q/bias']
    k = hs @ P['encode/cata/k/kernel'] + P['encode/cata/k/bias']
    v = hs @ P['encode/cata/v/kernel'] + P['encode/cata/v/bias']
    a = np.einsum('bhc,sbhc->bhs', q.reshape(b, hd, c), k.reshape(s, b, hd, c)) * h.dtype.type(c ** -0.5)
    pad = np.arange(s)[None, :] >= np.asarray(len_src)[:, None]            # (b,s)
    a = np.where(pad[:, None, :], -np.inf, a)
    a = np.exp(a - a.max(-1, keepdims=True))
    a = a / a.sum(-1, keepdims=True)
    y = np.einsum('bhs,sbhc->bhc', a, v.reshape(s, b, hd, c)).reshape(b, d)
    pp = y @ P['encode/cata/p/kernel'] + P['encode/cata/p/bias']
    x = h + pp
    mean = x.mean(-1, keepdims=True)
    var = ((x - mean) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + h.dtype.type(LN_EPS))
    xhat = (x - mean) * rstd
    out = xhat * P['encode/cata/LayerNorm/gamma'] + P['encode/cata/LayerNorm/beta']
    return out, dict(h=h, hs=hs, q=q, k=k, v=v, a=a, y=y, xhat=xhat, rstd=rstd)


def cata_backward(P, G, dout, cc):
    """gradient of cata_forward: accumulates the parameter gradients into G, returns (d h, d hs)."""
    h, hs, q, k, v, a, y, xhat, rstd = (cc[n] for n in ('h', 'hs', 'q', 'k', 'v', 'a', 'y', 'xhat', 'rstd'))
    s, b, d = hs.shape
    hd, c = ATT_HEADS, d // ATT_HEADS
    G['encode/cata/LayerNorm/gamma'] += (dout * xhat).sum(0)
    G['encode/cata/LayerNorm/beta'] += dout.sum(0)
    g = dout * P['encode/cata/LayerNorm/gamma']
    dx = rstd * (g - g.mean(-1, keepdims=True) - xhat * (g * xhat).mean(-1, keepdims=True))
    G['encode/cata/p/kernel'] += y.T @ dx
    G['encode/cata/p/bias'] += dx.sum(0)
    dy = (dx @ P['encode/cata/p/kernel'].T).reshape(b, hd, c)
    vh, kh, qh = v.reshape(s, b, hd, c), k.reshape(s, b, hd, c), q.reshape(b, hd, c)
    da = np.einsum('bhc,sbhc->bhs', dy, vh)
    dv = np.einsum('bhs,bhc->sbhc', a, dy).reshape(s, b, d)
    dsc = a * (da - (a * da).sum(-1, keepdims=True)) * h.dtype.type(c ** -0.5)
    dq = np.einsum('bhs,sbhc->bhc', dsc, kh).reshape(b, d)
    dk = np.einsum('bhs,bhc->sbhc', dsc, qh).reshape(s, b, d)
    G['encode/cata/q/kernel'] += h.T @ dq
    G['encode/cata/q/bias'] += dq.sum(0)
    hs2 = hs.reshape(s * b, d)
    G['encode/cata/k/kernel'] += hs2.T @ dk.reshape(s * b, d)
    G['encode/cata/k/bias'] += dk.sum((0, 1))
    G['encode/cata/v/kernel'] += hs2.T @ dv.reshape(s * b, d)
    G['encode/cata/v/bias'] += dv.sum((0, 1))
    dh = dx + dq @ P['encode/cata/q/kernel'].T
    dhs = dk @ P['encode/cata/k/kernel'].T + dv @ P['encode/cata/v/kernel'].T
    return dh, dhs


def gru_forward(x, h0, W, R, bW, bR):
    """x (T,b,in), h0 (b,H) -> hs (T,b,H), cache."""
    T, b, _ = x.shape
    H = R.shape[1]
    gx = x @ W.T + bW
    hs = np.empty((T, b, H), x.dtype)
    rs = np.empty_like(hs); us = np.empty_like(hs); ns = np.empty_like(hs); qs = np.empty_like(hs)
    h = h0
    for t in range(T):
        gh = h @ R.T + bR
        r = _sig(gx[t, :, :H] + gh[:, :H])
        u = _sig(gx[t, :, H:2 * H] + gh[:, H:2 * H])
        q = gh[:, 2 * H:]
        n = np.tanh(gx[t, :, 2 * H:] + r * q)
        h = (1 - u) * n + u * h
        hs[t] = h; rs[t] = r; us[t] = u; ns[t] = n; qs[t] = q
    return hs, (x, h0, hs, rs, us, ns, qs)


def gru_backward(dhs, dhT, cache, W, R):
    """dhs (T,b,H) grad wrt outputs, dhT (b,H) grad wrt the final state (or None).
    Returns dx, dh0, dW, dR, dbW, dbR."""
    x, h0, hs, rs, us, ns, qs = cache
    T, b, _ = x.shape
    H = R.shape[1]
    dgx = np.empty((T, b, 3 * H), x.dtype)
    dgh = np.empty((T, b, 3 * H), x.dtype)
    dh = np.zeros((b, H), x.dtype) if dhT is None else dhT.copy()
    for t in range(T - 1, -1, -1):
        hp = hs[t - 1] if t > 0 else h0
        d = dh + dhs[t]
        r, u, n, q = rs[t], us[t], ns[t], qs[t]
        dn = d * (1 - u) * (1 - n * n)
        du = d * (hp - n) * u * (1 - u)
        dr = dn * q * r * (1 - r)
        dgx[t, :, :H] = dr; dgx[t, :, H:2 * H] = du; dgx[t, :, 2 * H:] = dn
        dgh[t, :, :H] = dr; dgh[t, :, H:2 * H] = du; dgh[t, :, 2 * H:] = dn * r
        dh = d * u + dgh[t] @ R
    hprev = np.concatenate([h0[None], hs[:-1]], 0)
    dW = np.einsum('tbg,tbi->gi', dgx, x)
    dR = np.einsum('tbg,tbh->gh', dgh, hprev)
    return dgx @ W, dh, dW, dR, dgx.sum((0, 1)), dgh.sum((0, 1))


# --------------------------------------------------------------------------------------
# the graph: src/model.py:75-189
# --------------------------------------------------------------------------------------

def forward(P, cfg, src, tgt, mode='valid', step=0, keep=None, eps=None, want_prob=False, n_tokens_global=None, b_global=None):
    """P: canonical params (any float dtype).  src,tgt: int32 (b,T) eos-padded, batch-major.
    mode train: `keep` (t-1,b) 0/1 and `eps` (b,R) must be injected (TF's Philox streams
    cannot be reproduced).  Returns (Record-like dict, cache)."""
    assert mode in ('train', 'valid', 'infer')
    dt = P['embed/embedding'].dtype
    bos, eos = cfg.get('bos', 2), cfg.get('eos', 1)
    L, D = cfg.get('rnn_layers', 3), cfg.get('dim_emb', 512)
    H = D
    E = P['embed/embedding']
    o = dict(bos=bos, eos=eos, step=step)
    o.update(schedule(step, cfg.get('accelerate', 1e-4), cfg.get('learn_rate', 1e-3), dt))

    src_tm, msk_src, len_src = trim(np.ascontiguousarray(np.asarray(src, np.int32).T), eos)
    tgt_tm, not_eos_tgt, len_tgt = trim(np.ascontiguousarray(np.asarray(tgt, np.int32).T), eos)
    if mode == 'train':
        assert keep is not None and eps is not None, "train mode needs injected keep-mask and eps"
    lead, gold, msk_tgt = decoder_io(tgt_tm, not_eos_tgt, bos, eos, keep if mode == 'train' else None)
    o.update(lead=lead, gold=gold, msk_tgt=msk_tgt, len_src=len_src, len_tgt=len_tgt, src_tm=src_tm)
    b = src_tm.shape[1]

    emb_tgt = E[lead]
    x = E[src_tm]
    enc_caches = []
    stacks = enc_stacks(cfg)
    z0 = np.zeros((b, H), dt)
    if stacks is None:                                   # model.py:118-122 (config.json)
        for i in range(1, L + 1):
            pf, pb = 'encode/rnn%d/fwd/' % i, 'encode/rnn%d/bwd/' % i
            fwd, cf = gru_forward(x, z0, P[pf + 'W'], P[pf + 'R'], P[pf + 'bW'], P[pf + 'bR'])
            bwd, cb = gru_forward(reverse_sequence(x, len_src), z0, P[pb + 'W'], P[pb + 'R'], P[pb + 'bW'], P[pb + 'bR'])
            x = np.concatenate([fwd, reverse_sequence(bwd, len_src)], -1)
            enc_caches.append((cf, cb))
        hs = x
    else:                                                # model.py:124-131: L-layer stack(s), concatenated at the top only
        outs = []
        for k, st in enumerate(stacks):
            y = reverse_sequence(x, len_src) if k == 1 else x
            cs = []
            for j in range(L):
                p = '%sl%d/' % (st, j)
                y, c = gru_forward(y, z0, P[p + 'W'], P[p + 'R'], P[p + 'bW'], P[p + 'bR'])
                cs.append(c)
            enc_caches.append(cs)
            outs.append(reverse_sequence(y, len_src) if k == 1 else y)
        hs = np.concatenate(outs, -1)
    h = hs[len_src - 1, np.arange(b)]
    cata_cache = None
    if cfg.get('attentive', False):                      # model.py:136-145
        h, cata_cache = cata_forward(P, h, hs, len_src)
    mu = h @ P['latent/mu/kernel'] + P['latent/mu/bias']
    lv = h @ P['latent/lv/kernel'] + P['latent/lv/bias']
    z = mu
    if mode == 'train':
        z = mu + np.exp(0.5 * lv) * np.asarray(eps, dt)
    hx = z @ P['latent/ex/kernel'] + P['latent/ex/bias']
    o.update(mu=mu, lv=lv, z=z, enc_h=h)
    o['state_in'] = np.stack((hx,) * L)

    y = emb_tgt
    dec_caches = []
    state_ex = []
    for j in range(L):
        p = 'decode/rnn/l%d/' % j
        y, c = gru_forward(y, hx, P[p + 'W'], P[p + 'R'], P[p + 'bW'], P[p + 'bR'])
        dec_caches.append(c)
        state_ex.append(y[-1])
    o['state_ex'] = np.stack(state_ex)
    hd = y[msk_tgt] if mode != 'infer' else y.reshape(-1, H)
    ho = hd @ P['decode/out/kernel'] + P['decode/out/bias']
    scale = dt.type(D) ** dt.type(-0.5)
    if cfg.get('logit_use_embed', True):
        logits = ho @ (scale * E.T)                                          # model.py:165-166
    else:
        logits = ho @ P['logits/dense/kernel'] + P['logits/dense/bias']      # model.py:167-168
    o['logits'] = logits
    o['pred'] = np.argmax(logits, -1).astype(np.int32)
    if want_prob:
        e = np.exp(logits - logits.max(-1, keepdims=True))
        o['prob'] = e / e.sum(-1, keepdims=True)
    cache = None
    if mode != 'infer':
        labels = gold[msk_tgt]
        mx = logits.max(-1)
        lse = mx + np.log(np.exp(logits - mx[:, None]).sum(-1))
        gen_samp = lse - logits[np.arange(len(labels)), labels]
        kld_samp = 0.5 * (mu * mu + np.exp(lv) - lv - 1.0)
        o.update(labels=labels,
                 errt_samp=(labels != o['pred']).astype(np.float32), loss_gen_samp=gen_samp,
                 loss_kld_samp=kld_samp)
        n_norm = len(labels) if n_tokens_global is None else n_tokens_global
        k_norm = kld_samp.size if b_global is None else b_global * kld_samp.shape[1]
        o['errt'] = o['errt_samp'].sum() / n_norm
        o['loss_gen'] = gen_samp.sum() / n_norm
        o['loss_kld'] = kld_samp.sum() / k_norm
        o['loss'] = o['rate_anneal'] * o['loss_kld'] + o['loss_gen']
        cache = dict(E=E, lead=lead, src_tm=src_tm, len_src=len_src, msk_tgt=msk_tgt, labels=labels,
                     enc=enc_caches, dec=dec_caches, cata=cata_cache, h=h, mu=mu, lv=lv, z=z, eps=eps, hd=hd, ho=ho,
                     logits=logits, lse=lse, scale=scale, hs_shape=hs.shape, emb_tgt_shape=emb_tgt.shape,
                     mode=mode, anneal=o['rate_anneal'], n_norm=n_norm, k_norm=k_norm)
    return o, cache


def backward(P, cfg, cache):
    """Analytic gradient of `loss` (model.py:185) wrt every parameter (A18)."""
    L, D = cfg.get('rnn_layers', 3), cfg.get('dim_emb', 512)
    H = D
    E = cache['E']; dt = E.dtype
    G = {k: np.zeros_like(v) for k, v in P.items()}
    logits, labels, lse, scale = cache['logits'], cache['labels'], cache['lse'], cache['scale']
    N = len(labels)
    dlog = np.exp(logits - lse[:, None])
    dlog[np.arange(N), labels] -= 1.0
    dlog /= cache['n_norm']
    ho, hd = cache['ho'], cache['hd']
    if cfg.get('logit_use_embed', True):
        G['embed/embedding'] += scale * (dlog.T @ ho)
        dho = dlog @ (scale * E)
    else:
        G['logits/dense/kernel'] += ho.T @ dlog
        G['logits/dense/bias'] += dlog.sum(0)
        dho = dlog @ P['logits/dense/kernel'].T
    G['decode/out/kernel'] += hd.T @ dho
    G['decode/out/bias'] += dho.sum(0)
    dhd = dho @ P['decode/out/kernel'].T
    dy = np.zeros(cache['emb_tgt_shape'][:2] + (H,), dt)
    dy[cache['msk_tgt']] = dhd
    dhx = np.zeros_like(cache['dec'][0][1])
    for j in range(L - 1, -1, -1):
        p = 'decode/rnn/l%d/' % j
        dy, dh0, dW, dR, dbW, dbR = gru_backward(dy, None, cache['dec'][j], P[p + 'W'], P[p + 'R'])
        G[p + 'W'] += dW; G[p + 'R'] += dR; G[p + 'bW'] += dbW; G[p + 'bR'] += dbR
        dhx += dh0
    np.add.at(G['embed/embedding'], cache['lead'], dy)
    mu, lv, z = cache['mu'], cache['lv'], cache['z']
    G['latent/ex/kernel'] += z.T @ dhx
    G['latent/ex/bias'] += dhx.sum(0)
    dz = dhx @ P['latent/ex/kernel'].T
    a = cache['anneal'] / cache['k_norm']
    dmu = dz + a * mu
    dlv = a * 0.5 * (np.exp(lv) - 1.0)
    if cache['mode'] == 'train':
        dlv = dlv + dz * 0.5 * np.exp(0.5 * lv) * np.asarray(cache['eps'], dt)
    h = cache['h']
    G['latent/mu/kernel'] += h.T @ dmu; G['latent/mu/bias'] += dmu.sum(0)
    G['latent/lv/kernel'] += h.T @ dlv; G['latent/lv/bias'] += dlv.sum(0)
    dh = dmu @ P['latent/mu/kernel'].T + dlv @ P['latent/lv/kernel'].T
    len_src = cache['len_src']
    b = len(len_src)
    dhs = np.zeros(cache['hs_shape'], dt)
    if cache.get('cata') is not None:
        dh, dhs_att = cata_backward(P, G, dh, cache['cata'])
        # padded positions carry softmax weight 0, so d k and d v -- and with them dhs_att -- are exactly 0 there
        dhs += dhs_att
    dhs[len_src - 1, np.arange(b)] += dh
    stacks = enc_stacks(cfg)
    if stacks is None:
        for i in range(L, 0, -1):
            pf, pb = 'encode/rnn%d/fwd/' % i, 'encode/rnn%d/bwd/' % i
            cf, cb = cache['enc'][i - 1]
            dxf, _, dW, dR, dbW, dbR = gru_backward(dhs[..., :H], None, cf, P[pf + 'W'], P[pf + 'R'])
            G[pf + 'W'] += dW; G[pf + 'R'] += dR; G[pf + 'bW'] += dbW; G[pf + 'bR'] += dbR
            dxb, _, dW, dR, dbW, dbR = gru_backward(reverse_sequence(dhs[..., H:], len_src), None, cb, P[pb + 'W'], P[pb + 'R'])
            G[pb + 'W'] += dW; G[pb + 'R'] += dR; G[pb + 'bW'] += dbW; G[pb + 'bR'] += dbR
            dhs = dxf + reverse_sequence(dxb, len_src)
    else:
        dx = 0
        for k, st in enumerate(stacks):
            d = np.ascontiguousarray(dhs[..., k * H:(k + 1) * H])
            if k == 1:
                d = reverse_sequence(d, len_src)
            for j in range(L - 1, -1, -1):
                p = '%sl%d/' % (st, j)
                d, _, dW, dR, dbW, dbR = gru_backward(d, None, cache['enc'][k][j], P[p + 'W'], P[p + 'R'])
                G[p + 'W'] += dW; G[p + 'R'] += dR; G[p + 'bW'] += dbW; G[p + 'bR'] += dbR
            dx = dx + (reverse_sequence(d, len_src) if k == 1 else d)
        dhs = dx
    np.add.at(G['embed/embedding'], cache['src_tm'], dhs)
    return G


def adam_tf(P, G, M, Vv, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update #t (t counts from 1), in place.  A17."""
    dt = next(iter(P.values())).dtype.type
    lr_t = dt(lr) * np.sqrt(dt(1) - dt(beta2) ** dt(t)) / (dt(1) - dt(beta1) ** dt(t))
    for k in P:
        g = G[k]
        M[k] = dt(beta1) * M[k] + dt(1 - beta1) * g
        Vv[k] = dt(beta2) * Vv[k] + dt(1 - beta2) * g * g
        P[k] = P[k] - lr_t * M[k] / (np.sqrt(Vv[k]) + dt(eps))


def train_step(P, M, Vv, cfg, src, tgt, step, keep, eps):
    """One `sess.run(train_step)` (src/train.py:118): forward, backward, Adam, step+1."""
    o, cache = forward(P, cfg, src, tgt, 'train', step, keep, eps)
    G = backward(P, cfg, cache)
    adam_tf(P, G, M, Vv, step + 1, o['rate_update'])
    return o, G


def decode_greedy(P, cfg, z, steps=256):
    """src/model.py:204-219 -- greedy autoregressive decode, one step at a time."""
    L, D = cfg.get('rnn_layers', 3), cfg.get('dim_emb', 512)
    H = D
    bos, eos = cfg.get('bos', 2), cfg.get('eos', 1)
    E = P['embed/embedding']
    hx = z @ P['latent/ex/kernel'] + P['latent/ex/bias']
    s = [hx.copy() for _ in range(L)]
    x = np.full((len(z),), bos, np.int32)
    ys = []
    for _ in range(steps):
        y = E[x][None]
        for j in range(L):
            p = 'decode/rnn/l%d/' % j
            y, _ = gru_forward(y, s[j], P[p + 'W'], P[p + 'R'], P[p + 'bW'], P[p + 'bR'])
            s[j] = y[0]
        ho = y[0] @ P['decode/out/kernel'] + P['decode/out/bias']
        if cfg.get('logit_use_embed', True):
            x = np.argmax(ho @ (E.dtype.type(D) ** E.dtype.type(-0.5) * E.T), -1).astype(np.int32)
        else:
            x = np.argmax(ho @ P['logits/dense/kernel'] + P['logits/dense/bias'], -1).astype(np.int32)
        if np.all(x == eos):
            break
        ys.append(x[None])
    return np.concatenate(ys).T if ys else np.zeros((len(z), 0), np.int32)


# --------------------------------------------------------------------------------------
# synthetic workloads
# --------------------------------------------------------------------------------------

def synth_batch(b, kind='iac', vocab=8192, seed=0, eos=1, cap=None):
    """IAC-shaped: clip(rint(lognormal(ln 87, 1.0)),1,512); IBM-shaped: lognormal(ln 30,.5) cap 256;
    'full': every length == cap.  Zipf(1.0) token ids over 3..vocab-1 ('full': uniform)."""
    rng = np.random.default_rng(seed)
    if kind == 'iac':
        cap = cap or 512
        lens = np.clip(np.rint(rng.lognormal(np.log(87.0), 1.0, b)), 1, cap).astype(np.int64)
    elif kind == 'ibm':
        cap = cap or 256
        lens = np.clip(np.rint(rng.lognormal(np.log(30.0), 0.5, b)), 1, cap).astype(np.int64)
    elif kind == 'full':
        cap = cap or 512
        lens = np.full(b, cap, np.int64)
    else:
        raise ValueError(kind)
    nid = vocab - 3
    if kind == 'full':
        rows = [rng.integers(3, vocab, n).astype(np.int32) for n in lens]
    else:
        w = 1.0 / np.arange(1, nid + 1)
        cdf = np.cumsum(w / w.sum())
        rows = [(3 + np.minimum(np.searchsorted(cdf, rng.random(n)), nid - 1)).astype(np.int32) for n in lens]
    return vpack(rows, (b, int(lens.max())), eos, np.int32)
