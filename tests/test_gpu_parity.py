"""-m gpu: the CUDA path, called through the C ABI, against the CPU oracle (oracle/vae_oracle.py)
on the same seeded inputs.  fp32 validation mode: ELBO terms within 1e-3 relative (north_star);
index work bit exact.  The oracle is a restatement (TF cannot run here): parity is against it."""
import numpy as np
import pytest

from conftest import SMALL, ragged_batch
from oracle import vae_oracle as O

pytestmark = pytest.mark.gpu


def _mk(cfg, precision, seed=0, flags=0):
    from argsim_b200 import _lib
    h = _lib.Handle(precision=precision, flags=flags, **cfg)
    P = O.init_params(cfg, seed=seed, dtype=np.float64, bias_scale=0.1)
    h.set_params({k: v.astype(np.float32) for k, v in P.items()})
    return h, P


def _inject(cfg, tgt, seed):
    rng = np.random.default_rng(seed)
    keep = (rng.random(tgt.shape) < 0.7).astype(np.uint8)
    eps = rng.standard_normal((tgt.shape[0], cfg['dim_rep'])).astype(np.float32)
    return keep, eps


def _oracle_keep(keep, tgt, eos):
    """(b,T) batch-major mask -> the oracle's (t-1, b) time-major mask over the trimmed length"""
    tmax = int((tgt != eos).sum(1).max())
    return keep[:, :tmax].T


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def test_fp32_forward_losses_match_oracle():
    from argsim_b200 import _lib
    cfg = dict(SMALL)
    h, P = _mk(cfg, _lib.FP32_VALIDATE)
    src = ragged_batch(7, 13, cfg['dim_tgt'], 1)
    tgt = ragged_batch(7, 11, cfg['dim_tgt'], 2)
    o, _ = O.forward(P, cfg, src, tgt, 'valid', step=0)
    e = h.eval_step(src, tgt, want_pred=True)
    assert e['loss_gen_samp'].shape == o['loss_gen_samp'].shape
    np.testing.assert_allclose(e['loss_gen_samp'], o['loss_gen_samp'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(e['loss_kld_samp'], o['loss_kld_samp'], rtol=1e-4, atol=1e-6)
    np.testing.assert_array_equal(e['pred'], o['pred'])              # index work: bit exact
    np.testing.assert_array_equal(e['errt_samp'], o['errt_samp'])
    mu = h.embed(src)
    np.testing.assert_allclose(mu, o['mu'], rtol=1e-4, atol=1e-5)


def test_fp32_train_step_losses_and_grads_match_oracle():
    from argsim_b200 import _lib
    cfg = dict(SMALL)
    h, P = _mk(cfg, _lib.FP32_VALIDATE)
    src = ragged_batch(6, 12, cfg['dim_tgt'], 3)
    tgt = ragged_batch(6, 9, cfg['dim_tgt'], 4)
    keep, eps = _inject(cfg, tgt, 5)
    step = 12345
    h.step = step
    o, cache = O.forward(P, cfg, src, tgt, 'train', step=step, keep=_oracle_keep(keep, tgt, cfg['eos']), eps=eps.astype(np.float64))
    G = O.backward(P, cfg, cache)
    st = h.grad_step(src, tgt, keep=keep, eps=eps)
    assert st['n_tokens'] == len(o['labels'])
    assert rel(st['loss_gen'], o['loss_gen']) < 1e-3
    assert rel(st['loss_kld'], o['loss_kld']) < 1e-3
    assert rel(st['loss'], o['loss']) < 1e-3
    assert abs(st['errt'] - o['errt']) < 1e-6
    assert rel(st['rate_anneal'], o['rate_anneal']) < 1e-6 and rel(st['rate_update'], o['rate_update']) < 1e-6
    assert h.step == step  # grad_step does not advance
    for k in P:
        g = h.get_grad(k)
        scale = np.abs(G[k]).max() + 1e-12
        err = np.abs(g - G[k]).max() / scale
        assert err < 2e-4, (k, err)


def test_fp32_multi_step_adam_tracks_oracle():
    from argsim_b200 import _lib
    cfg = dict(SMALL)
    h, P = _mk(cfg, _lib.FP32_VALIDATE)
    M = {k: np.zeros_like(v) for k, v in P.items()}
    V = {k: np.zeros_like(v) for k, v in P.items()}
    for it in range(4):
        src = ragged_batch(5, 10, cfg['dim_tgt'], 10 + it)
        keep, eps = _inject(cfg, src, 20 + it)
        o, _ = O.train_step(P, M, V, cfg, src, src, it, _oracle_keep(keep, src, cfg['eos']), eps.astype(np.float64))
        st = h.train_step(src, src, keep=keep, eps=eps)
        assert st['step'] == it + 1
        for name, ref in (('loss', o['loss']), ('loss_gen', o['loss_gen']), ('loss_kld', o['loss_kld'])):
            assert rel(st[name], ref) < 1e-3, (it, name, st[name], ref)
    for k in P:
        p = h.get_param(k)
        assert np.abs(p - P[k]).max() < 2e-4, k   # 4 Adam steps move every weight by <= 4e-3
        m, v = h.get_opt_state(k)
        np.testing.assert_allclose(m, M[k], rtol=5e-3, atol=1e-7)


def test_edge_cases_len1_b1_and_contract_errors():
    from argsim_b200 import _lib
    cfg = dict(SMALL)
    h, P = _mk(cfg, _lib.FP32_VALIDATE)
    src = np.array([[7]], np.int32)               # b=1, len 1, no padding column at all
    o, _ = O.forward(P, cfg, src, src, 'valid')
    e = h.eval_step(src, src)
    np.testing.assert_allclose(e['loss_gen_samp'], o['loss_gen_samp'], rtol=1e-4, atol=1e-5)
    src = np.array([[7, 1, 1, 1], [9, 8, 7, 6], [5, 1, 1, 1]], np.int32)
    o, _ = O.forward(P, cfg, src, src, 'valid')
    e = h.eval_step(src, src)
    np.testing.assert_allclose(e['loss_gen_samp'], o['loss_gen_samp'], rtol=1e-4, atol=1e-5)
    with pytest.raises(RuntimeError):             # empty row: len_src must be >= 1 (model.py:135 would index -1)
        h.eval_step(np.array([[1, 1], [4, 5]], np.int32), np.array([[1, 1], [4, 5]], np.int32))
    with pytest.raises(RuntimeError):             # eos inside a sequence violates trim()'s contract
        h.eval_step(np.array([[4, 1, 5]], np.int32), np.array([[4, 1, 5]], np.int32))
    with pytest.raises(RuntimeError):             # id out of range
        h.eval_step(np.array([[4, 999]], np.int32), np.array([[4, 5]], np.int32))


def test_philox_keep_mask_matches_numpy_restatement():
    """library-drawn word dropout (no injected mask) == argsim_b200.rng.keep_mask fed back as injection"""
    from argsim_b200 import _lib, rng
    cfg = dict(SMALL)
    h, P = _mk(cfg, _lib.FP32_VALIDATE)
    h2, _ = _mk(cfg, _lib.FP32_VALIDATE)
    src = ragged_batch(6, 12, cfg['dim_tgt'], 7)
    eps = np.random.default_rng(0).standard_normal((6, cfg['dim_rep'])).astype(np.float32)
    for hh in (h, h2):
        hh.set_seed(1234)
        hh.step = 5000
    kw = _lib.schedule(5000, cfg['accelerate'], cfg['learn_rate'])['rate_keepwd']
    keep = rng.keep_mask(6, src.shape[1], kw, 1234, 5000)
    assert 0 < keep.mean() < 1
    a = h.grad_step(src, src, keep=None, eps=eps)
    b = h2.grad_step(src, src, keep=keep, eps=eps)
    assert a['loss_gen'] == b['loss_gen']


def test_decode_and_checkpoint_roundtrip(tmp_path):
    from argsim_b200 import _lib
    cfg = dict(SMALL)
    h, P = _mk(cfg, _lib.FP32_VALIDATE)
    z = np.random.default_rng(3).standard_normal((4, cfg['dim_rep'])).astype(np.float32)
    ref = O.decode_greedy({k: v.astype(np.float32) for k, v in P.items()}, cfg, z, steps=6)
    s = h.decode_init(z)
    x = np.full(4, cfg['bos'], np.int32)
    ys = []
    for _ in range(6):
        x, s = h.decode_step(x, s)
        if np.all(x == cfg['eos']):
            break
        ys.append(x)
    np.testing.assert_array_equal(np.stack(ys, 1), ref)
    # the same loop resident on the device (argsim_decode): identical tokens, the step budget is honoured, and the
    # all-eos stop test works (force eos by making its embedding row dominate the tied output projection)
    np.testing.assert_array_equal(h.decode(z, steps=6), ref)
    np.testing.assert_array_equal(h.decode(z, steps=3), ref[:, :3])
    long = O.decode_greedy({k: v.astype(np.float32) for k, v in P.items()}, cfg, z, steps=10)
    np.testing.assert_array_equal(h.decode(z, steps=10), long)
    P2 = {k: v.astype(np.float32).copy() for k, v in P.items()}
    P2['decode/out/kernel'][:] = 0.0
    P2['decode/out/bias'][:] = 0.05
    P2['embed/embedding'][:] = -np.abs(P2['embed/embedding'])
    P2['embed/embedding'][cfg['eos']] = 1.0          # logits = D^-1/2 * <out, E[v]>: eos wins for every input
    hz = _lib.Handle(precision=_lib.FP32_VALIDATE, **cfg)
    hz.set_params(P2)
    assert hz.decode(z, steps=20).shape == (4, 0)     # all-eos at the first step: nothing kept (the reference raises here)
    P2['embed/embedding'][7] = 2.0                    # token 7 wins forever: never stops, uses the whole budget
    hz.set_params(P2)
    out = hz.decode(z, steps=37)
    assert out.shape == (4, 37) and (out == 7).all()
    hz.close()
    src = ragged_batch(3, 8, cfg['dim_tgt'], 9)
    h.train_step(src, src)
    path = str(tmp_path / 'ckpt.bin')
    h.save(path)
    h2 = _lib.Handle(precision=_lib.FP32_VALIDATE, **cfg)
    h2.load(path)
    assert h2.step == 1
    for k in P:
        np.testing.assert_array_equal(h.get_param(k), h2.get_param(k))
        np.testing.assert_array_equal(h.get_opt_state(k)[1], h2.get_opt_state(k)[1])


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_untied_logit_projection_branch(prec):
    """logit_use_embed=false (src/model.py:167-168): logits = dense(h, dim_tgt) with its own (D,V) kernel and bias --
    losses, every gradient (the new kernel / bias, and the embedding that now only gets its gather parts), three Adam
    steps and the greedy decode against the oracle."""
    from argsim_b200 import _lib
    cfg = dict(SMALL, logit_use_embed=False)
    h, P = _mk(cfg, _lib.FP32_VALIDATE if prec == 'fp32' else _lib.BF16)
    assert h.param_shapes()['logits/dense/kernel'] == (cfg['dim_emb'], cfg['dim_tgt'])
    assert h.param_shapes()['logits/dense/bias'] == (cfg['dim_tgt'],)
    src = ragged_batch(7, 11, cfg['dim_tgt'], 70)
    tgt = ragged_batch(7, 9, cfg['dim_tgt'], 71)
    keep, eps = _inject(cfg, tgt, 72)
    h.step = 12000
    o, cache = O.forward(P, cfg, src, tgt, 'train', step=12000, keep=_oracle_keep(keep, tgt, cfg['eos']), eps=eps.astype(np.float64))
    G = O.backward(P, cfg, cache)
    st = h.grad_step(src, tgt, keep=keep, eps=eps)
    tol = 1e-3 if prec == 'fp32' else 1e-2
    for k in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(st[k], o[k]) < tol, (k, st[k], o[k])
    for k in P:
        g = h.get_grad(k).astype(np.float64)
        if prec == 'fp32':
            assert np.abs(g - G[k]).max() <= 2e-4 * np.abs(G[k]).max() + 1e-9, k
        elif np.linalg.norm(G[k]) > 1e-12:
            cos = (g.ravel() @ G[k].ravel()) / (np.linalg.norm(g) * np.linalg.norm(G[k]) + 1e-30)
            assert cos > 0.995, (k, cos)
    if prec == 'fp32':
        M = {k: np.zeros_like(v) for k, v in P.items()}
        V = {k: np.zeros_like(v) for k, v in P.items()}
        h.step = 0
        for it in range(3):
            oo, _ = O.train_step(P, M, V, cfg, src, tgt, it, _oracle_keep(keep, tgt, cfg['eos']), eps.astype(np.float64))
            s2 = h.train_step(src, tgt, keep=keep, eps=eps)
            assert rel(s2['loss'], oo['loss']) < 1e-3, (it, s2['loss'], oo['loss'])
        z = np.random.default_rng(4).standard_normal((3, cfg['dim_rep'])).astype(np.float32)
        ref = O.decode_greedy({k: v.astype(np.float32) for k, v in P.items()}, cfg, z, steps=6)
        np.testing.assert_array_equal(h.decode(z, steps=6), ref)
    h.close()


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('bidirectional,bidir_stacked', [(True, False), (False, True)])
def test_non_default_encoder_branches(prec, bidirectional, bidir_stacked):
    """src/model.py:124-131: the non-stacked bidirectional encoder (two L-layer stacks, concatenated at the top) and the
    unidirectional one (mu / lv read H instead of 2H columns) -- losses, all gradients, the embedding."""
    from argsim_b200 import _lib
    cfg = dict(SMALL, bidirectional=bidirectional, bidir_stacked=bidir_stacked)
    h, P = _mk(cfg, _lib.FP32_VALIDATE if prec == 'fp32' else _lib.BF16)
    assert set(h.param_shapes()) == set(P) and all(h.param_shapes()[k] == v.shape for k, v in P.items())
    src = ragged_batch(9, 13, cfg['dim_tgt'], 80)
    tgt = ragged_batch(9, 10, cfg['dim_tgt'], 81)
    keep, eps = _inject(cfg, tgt, 82)
    h.step = 9000
    o, cache = O.forward(P, cfg, src, tgt, 'train', step=9000, keep=_oracle_keep(keep, tgt, cfg['eos']), eps=eps.astype(np.float64))
    G = O.backward(P, cfg, cache)
    st = h.grad_step(src, tgt, keep=keep, eps=eps)
    tol = 1e-3 if prec == 'fp32' else 1e-2
    for k in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(st[k], o[k]) < tol, (k, st[k], o[k])
    for k in P:
        g = h.get_grad(k).astype(np.float64)
        if prec == 'fp32':
            assert np.abs(g - G[k]).max() <= 2e-4 * np.abs(G[k]).max() + 1e-9, k
        elif np.linalg.norm(G[k]) > 1e-12:
            cos = (g.ravel() @ G[k].ravel()) / (np.linalg.norm(g) * np.linalg.norm(G[k]) + 1e-30)
            assert cos > 0.99, (k, cos)
    mu = h.embed(src)
    assert np.abs(mu - o['mu']).max() <= (1e-4 if prec == 'fp32' else 3e-2) * np.abs(o['mu']).max()
    h.close()


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('bidirectional,bidir_stacked', [(True, True), (True, False), (False, True)])
def test_attentive_branch(prec, bidirectional, bidir_stacked):
    """src/model.py:136-145 (attentive=true; repaired form stated at oracle/vae_oracle.py:cata_forward): the final state
    attends over its own sequence's outputs, residual + layer norm -- losses, every gradient (the q/k/v/p affines, the
    layer-norm scale and shift, and the encoder underneath, which now receives gradient at every step), the embedding,
    and two Adam steps."""
    from argsim_b200 import _lib
    cfg = dict(SMALL, bidirectional=bidirectional, bidir_stacked=bidir_stacked, attentive=True)
    h, P = _mk(cfg, _lib.FP32_VALIDATE if prec == 'fp32' else _lib.BF16)
    assert set(h.param_shapes()) == set(P) and all(h.param_shapes()[k] == v.shape for k, v in P.items())
    src = ragged_batch(9, 13, cfg['dim_tgt'], 90)
    tgt = ragged_batch(9, 10, cfg['dim_tgt'], 91)
    keep, eps = _inject(cfg, tgt, 92)
    h.step = 9000
    o, cache = O.forward(P, cfg, src, tgt, 'train', step=9000, keep=_oracle_keep(keep, tgt, cfg['eos']), eps=eps.astype(np.float64))
    G = O.backward(P, cfg, cache)
    st = h.grad_step(src, tgt, keep=keep, eps=eps)
    tol = 1e-3 if prec == 'fp32' else 1e-2
    for k in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(st[k], o[k]) < tol, (k, st[k], o[k])
    gmax = max(np.abs(v).max() for v in G.values())
    for k in P:
        g = h.get_grad(k).astype(np.float64)
        if k == 'encode/cata/k/bias':   # softmax is shift invariant: the gradient is identically zero
            assert np.abs(g).max() <= (1e-5 if prec == 'fp32' else 1e-2) * gmax, (k, np.abs(g).max())
        elif prec == 'fp32':
            assert np.abs(g - G[k]).max() <= 2e-4 * np.abs(G[k]).max() + 1e-9, k
        elif np.linalg.norm(G[k]) > 1e-12:
            cos = (g.ravel() @ G[k].ravel()) / (np.linalg.norm(g) * np.linalg.norm(G[k]) + 1e-30)
            assert cos > 0.99, (k, cos)
    mu = h.embed(src)
    assert np.abs(mu - o['mu']).max() <= (1e-4 if prec == 'fp32' else 3e-2) * np.abs(o['mu']).max()
    if prec == 'fp32':   # two optimizer steps track the oracle's parameters (the new tensors included)
        Po = {k: v.copy() for k, v in P.items()}
        M = {k: np.zeros_like(v) for k, v in P.items()}
        Vv = {k: np.zeros_like(v) for k, v in P.items()}
        h.step = 9000
        for i in range(2):
            O.train_step(Po, M, Vv, cfg, src, tgt, 9000 + i, _oracle_keep(keep, tgt, cfg['eos']), eps.astype(np.float64))
            h.train_step(src, tgt, keep=keep, eps=eps)
        for k in ('encode/cata/q/kernel', 'encode/cata/v/kernel', 'encode/cata/LayerNorm/gamma', 'encode/cata/p/bias'):
            assert np.abs(h.get_param(k) - Po[k]).max() <= 2e-4, k
    h.close()


def test_attentive_embedding_across_micro_batches():
    """argsim_embed with attentive=true on more rows than one micro-batch holds (the attention reads each micro-batch's own
    packed layout) and on a sequence long enough for the large score buffer: sampled rows against the oracle"""
    from argsim_b200 import _lib
    cfg = dict(SMALL, attentive=True)
    h, P = _mk(cfg, _lib.FP32_VALIDATE)
    src = ragged_batch(1100, 21, cfg['dim_tgt'], 95)
    mu = h.embed(src)
    pick = np.random.default_rng(96).choice(len(src), 24, replace=False)
    want = O.forward(P, cfg, src[pick], src[pick], 'infer')[0]['mu']
    assert np.abs(mu[pick] - want).max() <= 2e-4 * np.abs(want).max()
    long = ragged_batch(3, 13000, cfg['dim_tgt'], 97, tmin=12500)
    mu = h.embed(long)
    want = O.forward(P, cfg, long, long, 'infer')[0]['mu']
    assert np.abs(mu - want).max() <= 5e-4 * np.abs(want).max()
    h.close()


def test_tf_bundle_export_import_roundtrip(tmp_path):
    """Saver.save(tf_format=True) -> <path>.index / .data-00000-of-00001 under the reference's variable names -> Saver.restore
    into a fresh session: parameters (up to the r/u bias split the cuDNN-canonical form cannot keep), Adam slots, the
    step and therefore the next training step are reproduced."""
    from argsim_b200 import _lib, model as M
    cfg = dict(SMALL)
    M.reset() if hasattr(M, 'reset') else M._state.update(config=None, session=None)
    m = M.vAe('valid', **cfg)
    sess = M.Session(precision='fp32')
    M.global_variables_initializer(sess)
    src = ragged_batch(6, 9, cfg['dim_tgt'], 5)
    keep, eps = _inject(cfg, src, 6)
    for _ in range(3):
        sess.handle.train_step(src, src, keep=keep, eps=eps)
    path = str(tmp_path / 'master0')
    M.Saver().save(sess, path, tf_format=True)
    assert (tmp_path / 'master0.index').exists() and (tmp_path / 'master0.data-00000-of-00001').exists()
    ref = sess.handle.train_step(src, src, keep=keep, eps=eps)
    sess.close()
    M._state.update(config=None, session=None)
    m = M.vAe('valid', **cfg)
    sess2 = M.Session(precision='fp32')
    sv = M.Saver()
    sv.restore(sess2, path)
    assert sv.unplaced == []
    assert sess2.handle.step == 3
    got = sess2.handle.train_step(src, src, keep=keep, eps=eps)
    for k in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(got[k], ref[k]) < 1e-5, (k, got[k], ref[k])
    sess2.close()
    M._state.update(config=None, session=None)
