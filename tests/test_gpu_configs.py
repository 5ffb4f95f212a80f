"""-m gpu: every BASELINE.json config at its OWN dimensions against the oracle ("next round" item 1).

  * configs[1] (C1: config.json model, seed-0 IAC-shaped batch of 64, S = 8,808, N = 8,872, longest row 512): ALL
    gradients of one step against torch autograd of the oracle twin (oracle/vae_torch.py) -- fp32 validation mode within
    2e-4 of each tensor's max, bf16 mode cosine > 0.995 and norm ratio within 3 % -- and a 10-step bf16 training
    trajectory (injected keep masks and eps, TF-form Adam on both sides): ELBO / reconstruction / KL within 1e-2 relative
    at every step.  This is the test that puts the machinery the benchmark runs (171-step encoder segment chains, 64-step
    decoder wavefront, slice budget, padded grids, side stream, early Adam) against the oracle rather than against
    another CUDA path;
  * configs[4] dimensions (dim_emb 2048, dim_tgt 32768, dim_rep 4096: generic recurrence with the tcgen05 GEMM streaming
    R, split-K reduce-add, chunked logits) on a tiny batch, both precisions;
  * configs[3]: a batch of 4096 IBM-shaped rows through argsim_embed (micro-batches of 512), 64 sampled rows against
    the oracle's mu;
  * A22 (src/model.py:8-15,109): argsim_init_params read back -- bounds, zero biases, per-gate scaling;
  * with the rows' GLOBAL indices given (argsim_set_global_rows) the un-injected Philox randomness makes a
    batch dealt over two "ranks" equal to the same batch on one.

The oracle is a restatement (TensorFlow cannot run here): parity is against it -- "parity unpinned", DESIGN.md section 2.
"""
import numpy as np
import pytest

from oracle import vae_oracle as O
from test_gpu_parity import _oracle_keep, rel

pytestmark = pytest.mark.gpu
C1 = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
STEP = 15000   # rate 1.5: keep 0.82, anneal 0.905 -- both loss terms matter


def _torch():
    import torch
    from oracle import vae_torch as T
    return torch, T


@pytest.fixture(scope='module')
def c1():
    from argsim_b200.synth import synth_batch
    src = synth_batch(64, 'iac', C1['dim_tgt'], seed=0)
    assert int((src != 1).sum()) == 8808 and src.shape[1] == 512
    P = O.init_params(C1, seed=0, dtype=np.float32, bias_scale=0.05)
    return dict(src=src, P=P)


def _draw(src, seed, keep_rate=0.82):
    rng = np.random.default_rng(seed)
    keep = (rng.random(src.shape) < keep_rate).astype(np.uint8)
    eps = rng.standard_normal((src.shape[0], C1['dim_rep'])).astype(np.float32)
    return keep, eps


def test_c1_full_size_all_gradients_match_torch_autograd(c1):
    from argsim_b200 import _lib
    torch, T = _torch()
    src, P = c1['src'], c1['P']
    keep, eps = _draw(src, 5)
    Pt = T.to_torch(P, torch.float32, requires_grad=True)
    o = T.forward(Pt, C1, src, src, 'train', step=STEP, keep=_oracle_keep(keep, src, 1).astype(np.int64), eps=eps)
    o['loss'].backward()
    G = {k: v.grad.numpy() for k, v in Pt.items()}
    for name, prec in (('fp32', _lib.FP32_VALIDATE), ('bf16', _lib.BF16)):
        h = _lib.Handle(precision=prec, **C1)
        h.set_params(P)
        h.step = STEP
        st = h.grad_step(src, src, keep=keep, eps=eps)
        assert st['n_tokens'] == 8872
        tol = 1e-3 if name == 'fp32' else 1e-2
        for k in ('loss', 'loss_gen', 'loss_kld'):
            assert rel(st[k], float(o[k])) < tol, (name, k, st[k], float(o[k]))
        for k in P:
            g, r = h.get_grad(k).astype(np.float64).ravel(), G[k].astype(np.float64).ravel()
            if name == 'fp32':
                err = np.abs(g - r).max() / (np.abs(r).max() + 1e-30)
                assert err < 2e-4, (name, k, err)
            elif np.linalg.norm(r) < 1e-12:
                # exactly zero in the oracle too: the top layer's reverse GRU feeds the latent only through its output at
                # t = len-1 (src/model.py:135), its FIRST step, where h_prev = 0 -- so d R of encode/rnn3/bwd vanishes
                assert np.linalg.norm(g) < 1e-9, (name, k, np.linalg.norm(g))
            else:
                cos = float(g @ r) / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
                ratio = np.linalg.norm(g) / (np.linalg.norm(r) + 1e-30)
                assert cos > 0.995, (name, k, cos)
                assert abs(ratio - 1.0) < 0.03, (name, k, ratio)
        h.close()


def test_attentive_at_config_dims_matches_torch_autograd():
    """attentive=true (src/model.py:136-145, repaired form: oracle/vae_oracle.py:cata_forward) at config.json's dimensions
    (8 heads of 128 over the 1024-wide encoder output): every gradient against torch autograd, both precisions; the
    persistent recurrence kernels now receive d hs at every step of the top layer, not only at the final state"""
    from argsim_b200 import _lib
    from argsim_b200.synth import synth_batch
    torch, T = _torch()
    cfg = dict(C1, attentive=True)
    src = synth_batch(64, 'iac', cfg['dim_tgt'], seed=3)[:24, :96]
    src[:, -1] = 1                      # every row keeps an eos pad after the cut
    P = O.init_params(cfg, seed=2, dtype=np.float32, bias_scale=0.05)
    keep, eps = _draw(src, 6)
    Pt = T.to_torch(P, torch.float32, requires_grad=True)
    o = T.forward(Pt, cfg, src, src, 'train', step=STEP, keep=_oracle_keep(keep, src, 1).astype(np.int64), eps=eps)
    o['loss'].backward()
    G = {k: v.grad.numpy() for k, v in Pt.items()}
    gmax = max(float(np.abs(v).max()) for v in G.values())
    assert np.linalg.norm(G['encode/rnn3/bwd/R']) > 1e-9        # no longer the exactly-zero case of the default graph
    for name, prec in (('fp32', _lib.FP32_VALIDATE), ('bf16', _lib.BF16)):
        h = _lib.Handle(precision=prec, **cfg)
        h.set_params(P)
        h.step = STEP
        st = h.grad_step(src, src, keep=keep, eps=eps)
        tol = 1e-3 if name == 'fp32' else 1e-2
        for k in ('loss', 'loss_gen', 'loss_kld'):
            assert rel(st[k], float(o[k])) < tol, (name, k, st[k], float(o[k]))
        for k in P:
            g, r = h.get_grad(k).astype(np.float64).ravel(), G[k].astype(np.float64).ravel()
            if k == 'encode/cata/k/bias':   # identically zero (softmax shift invariance); fp32 rounding in the oracle too
                assert np.abs(g).max() < (1e-5 if name == 'fp32' else 1e-2) * gmax, (name, k, np.abs(g).max())
            elif name == 'fp32':
                err = np.abs(g - r).max() / (np.abs(r).max() + 1e-30)
                assert err < 5e-4, (name, k, err)
            else:
                cos = float(g @ r) / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
                ratio = np.linalg.norm(g) / (np.linalg.norm(r) + 1e-30)
                assert cos > 0.99, (name, k, cos)
                assert abs(ratio - 1.0) < 0.05, (name, k, ratio)
        mu = h.embed(src)
        want = T.forward(T.to_torch(P), cfg, src, src, 'infer', encoder_only=True)['mu'].numpy()
        assert np.abs(mu - want).max() <= (2e-4 if name == 'fp32' else 3e-2) * np.abs(want).max(), name
        h.close()


def test_c1_bf16_ten_step_trajectory_matches_oracle(c1):
    """10 training steps (forward, backward, TF-form Adam) on both sides with the same injected randomness."""
    from argsim_b200 import _lib
    torch, T = _torch()
    src, P = c1['src'], c1['P']
    Pt = T.to_torch(P, torch.float32, requires_grad=True)
    M = {k: torch.zeros_like(v) for k, v in Pt.items()}
    V = {k: torch.zeros_like(v) for k, v in Pt.items()}
    h = _lib.Handle(precision=_lib.BF16, **C1)
    h.set_params(P)
    h.step = STEP
    worst = 0.0
    import os
    for i in range(int(os.environ.get('ARGSIM_TRAJ_STEPS', 10))):
        keep, eps = _draw(src, 100 + i)
        o = T.train_step(Pt, M, V, C1, src, src, STEP + i, _oracle_keep(keep, src, 1).astype(np.int64), eps)
        st = h.train_step(src, src, keep=keep, eps=eps)
        assert st['step'] == STEP + i + 1
        for k in ('loss', 'loss_gen', 'loss_kld'):
            r = rel(st[k], float(o[k]))
            worst = max(worst, r)
            assert r < 1e-2, (i, k, st[k], float(o[k]), r)
    # the weights moved the same way: after 10 Adam steps the largest tensors still agree
    for k in ('embed/embedding', 'encode/rnn1/fwd/R', 'decode/rnn/l2/W', 'latent/mu/kernel'):
        a, b = h.get_param(k).ravel().astype(np.float64), Pt[k].detach().numpy().ravel().astype(np.float64)
        d0 = P[k].ravel().astype(np.float64)
        cos = float((a - d0) @ (b - d0)) / (np.linalg.norm(a - d0) * np.linalg.norm(b - d0) + 1e-30)
        assert cos > 0.9, (k, cos)     # update directions (Adam is sign-like on small gradients: bf16 flips some)
    h.close()
    print('worst relative deviation over 10 bf16 steps: %.2e' % worst)


SCALED = dict(dim_tgt=32768, dim_emb=2048, dim_rep=4096, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)


def test_scaled_config_dims_match_oracle():
    """BASELINE configs[4] dimensions, tiny batch (b = 4, T <= 12): losses and a few gradients, both precisions."""
    from argsim_b200 import _lib
    from conftest import ragged_batch
    torch, T = _torch()
    cfg = SCALED
    P = O.init_params(cfg, seed=1, dtype=np.float32, bias_scale=0.05)
    src = ragged_batch(4, 12, cfg['dim_tgt'], 11)
    tgt = ragged_batch(4, 10, cfg['dim_tgt'], 12)
    rng = np.random.default_rng(13)
    keep = (rng.random(tgt.shape) < 0.8).astype(np.uint8)
    eps = rng.standard_normal((4, cfg['dim_rep'])).astype(np.float32)
    Pt = T.to_torch(P, torch.float32, requires_grad=True)
    o = T.forward(Pt, cfg, src, tgt, 'train', step=STEP, keep=_oracle_keep(keep, tgt, 1).astype(np.int64), eps=eps)
    o['loss'].backward()
    check = ('decode/out/kernel', 'decode/rnn/l0/R', 'decode/rnn/l2/W', 'latent/ex/kernel', 'latent/mu/kernel',
             'encode/rnn3/bwd/R', 'encode/rnn2/fwd/W', 'encode/rnn1/fwd/bR', 'embed/embedding')
    G = {k: Pt[k].grad.numpy() for k in check}
    mu_ref = o['mu'].detach().numpy()
    for name, prec in (('fp32', _lib.FP32_VALIDATE), ('bf16', _lib.BF16)):
        h = _lib.Handle(precision=prec, **cfg)
        h.set_params(P)
        h.step = STEP
        st = h.grad_step(src, tgt, keep=keep, eps=eps)
        assert st['n_tokens'] == int(((tgt != 1).sum(1) + 1).sum())
        tol = 1e-3 if name == 'fp32' else 1e-2
        for k in ('loss', 'loss_gen', 'loss_kld'):
            assert rel(st[k], float(o[k])) < tol, (name, k, st[k], float(o[k]))
        for k in check:
            g, r = h.get_grad(k).astype(np.float64).ravel(), G[k].astype(np.float64).ravel()
            if name == 'fp32':
                err = np.abs(g - r).max() / (np.abs(r).max() + 1e-30)
                assert err < 2e-4, (name, k, err)
            elif np.linalg.norm(r) < 1e-12:
                assert np.linalg.norm(g) < 1e-9, (name, k, np.linalg.norm(g))
            else:
                cos = float(g @ r) / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
                assert cos > 0.995, (name, k, cos)
        mu = h.embed(src)
        scale = np.abs(mu_ref).max()
        assert np.abs(mu - mu_ref).max() <= (2e-4 if name == 'fp32' else 3e-2) * scale, (name, np.abs(mu - mu_ref).max(), scale)
        h.close()


def test_embed_batch_4096_ibm_shaped_matches_oracle(c1):
    """BASELINE configs[3]: encoder-only mu of 4096 IBM-shaped rows (argsim_embed -> micro-batches of 512 rows);
    64 sampled rows against the oracle (rows are independent in 'infer' mode, so the oracle runs on those 64 alone)."""
    from argsim_b200 import _lib
    from argsim_b200.synth import synth_batch
    torch, T = _torch()
    P = c1['P']
    data = synth_batch(4096, 'ibm', C1['dim_tgt'], seed=0)
    pick = np.sort(np.random.default_rng(3).choice(4096, 64, replace=False))
    pick[0], pick[-1] = 0, 4095
    lens = (data != 1).sum(1)
    pick[1] = int(np.argmax(lens))                      # the longest row and the shortest one are in the sample
    pick[2] = int(np.argmin(lens))
    sub = data[pick]
    sub = sub[:, :int((sub != 1).sum(1).max())]
    with torch.no_grad():
        ref = T.forward(T.to_torch(P, torch.float32), C1, sub, sub, encoder_only=True)['mu'].numpy()
    scale = np.abs(ref).max()
    for name, prec, tol in (('fp32', _lib.FP32_VALIDATE, 2e-4), ('bf16', _lib.BF16, 3e-2)):
        h = _lib.Handle(precision=prec, **C1)
        h.set_params(P)
        mu = h.embed(data)
        assert mu.shape == (4096, 1024) and np.isfinite(mu).all()
        err = np.abs(mu[pick] - ref).max()
        assert err <= tol * scale, (name, err, scale)
        if name == 'bf16':
            cos = (mu[pick] * ref).sum(1) / (np.linalg.norm(mu[pick], axis=1) * np.linalg.norm(ref, axis=1))
            assert cos.min() > 0.999, cos.min()
        h.close()


@pytest.mark.parametrize('kind', ['stacked', 'untied_unidirectional'])
def test_init_params_read_back(kind):
    """A22: src/model.py:8-15 (variance_scaling(1, fan_avg, uniform) per (H, in) gate sub-matrix, zero biases),
    src/model.py:109 (embedding uniform +-sqrt(6 / (V/D + 1))), tf.layers.dense default glorot_uniform."""
    from argsim_b200 import _lib
    cfg = dict(dim_tgt=1024, dim_emb=128, dim_rep=192, rnn_layers=2, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
    if kind != 'stacked':
        cfg.update(bidirectional=False, bidir_stacked=False, logit_use_embed=False)
    h = _lib.Handle(precision=_lib.FP32_VALIDATE, **cfg)
    h.init_params(7)
    V, D, H = cfg['dim_tgt'], cfg['dim_emb'], cfg['dim_emb']
    seen = set()
    for name, shape in h.param_shapes().items():
        w = h.get_param(name)
        assert w.shape == shape and np.isfinite(w).all()
        if len(shape) == 1:
            assert not w.any(), name                              # every bias starts at zero
            continue
        if name == 'embed/embedding':
            bound = np.sqrt(6.0 / (V / D + 1.0))
        elif name.endswith('/W') or name.endswith('/R'):
            assert shape[0] == 3 * H
            bound = np.sqrt(6.0 / (shape[1] + H))                 # fan_avg of ONE (H, in) gate sub-matrix, not of (3H, in)
            for g in range(3):                                    # every gate fills its own range
                assert np.abs(w[g * H:(g + 1) * H]).max() > 0.97 * bound, (name, g)
        else:
            bound = np.sqrt(6.0 / (shape[0] + shape[1]))          # glorot_uniform of tf.layers.dense
        assert np.abs(w).max() <= bound * (1 + 1e-6), (name, np.abs(w).max(), bound)
        assert np.abs(w).max() > 0.97 * bound, (name, np.abs(w).max(), bound)
        assert abs(w.mean()) < 0.05 * bound, name
        assert abs(w.std() - bound / np.sqrt(3.0)) < 0.05 * bound, (name, w.std(), bound / np.sqrt(3.0))   # uniform: std = bound / sqrt 3
        seen.add(name.split('/')[0])
    assert seen >= {'embed', 'encode', 'latent', 'decode'}
    a = h.get_param('embed/embedding')
    h.init_params(8)
    assert np.abs(h.get_param('embed/embedding') - a).max() > 0    # the seed matters
    h.init_params(7)
    np.testing.assert_array_equal(h.get_param('embed/embedding'), a)    # and reproduces
    assert h.step == 0
    h.close()


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_global_row_keying_makes_shards_equal_the_full_batch(mode):
    """No injection: word dropout and eps come from the library's Philox streams keyed by (seed, step, GLOBAL row,
    position).  One handle plays both ranks of a 2-way data-parallel step in turn (rows dealt by parallel.shard_batch,
    global normalisers): the shards' loss terms and gradients add up to the full batch's."""
    from argsim_b200 import _lib, parallel, rng
    from conftest import SMALL, ragged_batch
    cfg = dict(SMALL)
    prec = _lib.FP32_VALIDATE if mode == 'fp32' else _lib.BF16
    h = _lib.Handle(precision=prec, **cfg)
    P = O.init_params(cfg, seed=0, dtype=np.float32, bias_scale=0.1)
    h.set_params(P)
    h.set_seed(99)
    h.step = 7000
    src = ragged_batch(9, 14, cfg['dim_tgt'], 21)
    tgt = ragged_batch(9, 11, cfg['dim_tgt'], 22)
    full = h.grad_step(src, tgt, rows=np.arange(9))
    Gf = {k: h.get_grad(k).astype(np.float64) for k in P}
    # the keep mask the library drew is the documented stream (argsim_b200/rng.py) of the global rows
    keep = rng.keep_mask(9, tgt.shape[1], full['rate_keepwd'], 99, 7000, rows=np.arange(9))
    eps_free = h.grad_step(src, tgt, keep=keep, rows=np.arange(9))
    assert eps_free['loss_gen'] == full['loss_gen'] and eps_free['loss_kld'] == full['loss_kld']
    acc = dict(loss_gen=0.0, loss_kld=0.0, errt=0.0)
    Gs = {k: np.zeros_like(v) for k, v in Gf.items()}
    for rank in range(2):
        s, t, rows, n_glob, b_glob = parallel.shard_batch(src, tgt, 2, rank)
        st = h.grad_step(s, t, n_tokens_global=n_glob, b_global=b_glob, rows=rows)
        for k in acc:
            acc[k] += st[k]
        for k in Gs:
            Gs[k] += h.get_grad(k)
    tol = 2e-5 if mode == 'fp32' else 2e-2
    for k in acc:
        assert rel(acc[k], full[k]) < tol or abs(acc[k] - full[k]) < 1e-6, (k, acc[k], full[k])
    for k in Gs:
        err = np.abs(Gs[k] - Gf[k]).max() / (np.abs(Gf[k]).max() + 1e-30)
        assert err < (1e-4 if mode == 'fp32' else 5e-2), (k, err)
    # keyed by a rank-local offset instead, the same shards see other streams (what round 1 did)
    s, t, rows, n_glob, b_glob = parallel.shard_batch(src, tgt, 2, 1)
    st_local = h.grad_step(s, t, n_tokens_global=n_glob, b_global=b_glob, row0=5)
    st_glob = h.grad_step(s, t, n_tokens_global=n_glob, b_global=b_glob, rows=rows)
    assert st_local['loss_gen'] != st_glob['loss_gen']
    with pytest.raises(RuntimeError):
        h.grad_step(s, t, rows=np.arange(len(s)) - 1)     # negative index
    h.close()
