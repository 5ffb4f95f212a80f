"""-m gpu: the hot path at BASELINE.json's FULL size (configs[1]: config.json model, the seed-0 IAC-shaped batch of 64
sequences, S = 8,808 source tokens, N = 8,872 target rows, longest row 512) checked through size-independent
properties the domain offers, plus one oracle comparison at that size:

  * the fp32 validation mode against the torch-CPU twin of the oracle (oracle/vae_torch.py, ~10 s): ELBO /
    reconstruction / KL within 1e-3 relative (north_star), token counts exact;
  * the bf16 mode against the fp32 mode on the same weights and injected randomness: within 1e-2 relative;
  * index pipeline: eos padding beyond the longest row and a permutation of the batch rows do not change the step
    (counts bit exact; sums within fp32 reassociation);
  * valid-mode per-sample outputs average to the batch terms (src/train.py:104-113 relies on that);
  * embedding (src/model.py:194-201): a row's mu does not depend on its batch neighbours; micro-batching is invisible.
"""
import numpy as np
import pytest

from oracle import vae_oracle as O
from test_gpu_parity import _oracle_keep, rel

pytestmark = pytest.mark.gpu
CFG = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
STEP = 15000   # rate 1.5: keep 0.82, anneal 0.905 -- both loss terms matter


@pytest.fixture(scope='module')
def world():
    from argsim_b200 import _lib
    from argsim_b200.synth import synth_batch
    src = synth_batch(64, 'iac', CFG['dim_tgt'], seed=0)
    assert int((src != 1).sum()) == 8808 and src.shape[1] == 512
    P = O.init_params(CFG, seed=0, dtype=np.float32, bias_scale=0.05)
    rng = np.random.default_rng(5)
    keep = (rng.random(src.shape) < 0.82).astype(np.uint8)
    eps = rng.standard_normal((64, CFG['dim_rep'])).astype(np.float32)
    hs = {}
    for name, prec in (('fp32', _lib.FP32_VALIDATE), ('bf16', _lib.BF16)):
        h = _lib.Handle(precision=prec, **CFG)
        h.set_params(P)
        h.step = STEP
        hs[name] = h
    yield dict(src=src, P=P, keep=keep, eps=eps, h=hs)
    for h in hs.values():
        h.close()


def test_fp32_full_size_matches_torch_oracle(world):
    import torch
    from oracle import vae_torch as T
    w = world
    torch.set_num_threads(max(1, torch.get_num_threads()))
    Pt = T.to_torch(w['P'], torch.float32)
    with torch.no_grad():
        o = T.forward(Pt, CFG, w['src'], w['src'], 'train', step=STEP, keep=_oracle_keep(w['keep'], w['src'], 1).astype(np.int64),
                      eps=w['eps'])
    st = w['h']['fp32'].grad_step(w['src'], w['src'], keep=w['keep'], eps=w['eps'])
    assert st['n_tokens'] == 8872
    for k in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(st[k], float(o[k])) < 1e-3, (k, st[k], float(o[k]))


def test_bf16_full_size_within_1e2_of_fp32(world):
    w = world
    a = w['h']['fp32'].grad_step(w['src'], w['src'], keep=w['keep'], eps=w['eps'])
    b = w['h']['bf16'].grad_step(w['src'], w['src'], keep=w['keep'], eps=w['eps'])
    assert a['n_tokens'] == b['n_tokens'] == 8872
    for k in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(b[k], a[k]) < 1e-2, (k, b[k], a[k])
    assert abs(a['errt'] - b['errt']) < 0.02


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_padding_and_row_permutation_invariance(world, mode):
    w = world
    h = w['h'][mode]
    base = h.grad_step(w['src'], w['src'], keep=w['keep'], eps=w['eps'])
    # 37 extra eos columns: trim (src/util_tf.py:54-57) removes them before anything is computed
    pad = np.full((64, 37), 1, np.int32)
    src2 = np.concatenate([w['src'], pad], 1)
    keep2 = np.concatenate([w['keep'], np.ones((64, 37), np.uint8)], 1)
    st = h.grad_step(src2, src2, keep=keep2, eps=w['eps'])
    assert st['n_tokens'] == base['n_tokens']
    for k in ('loss', 'loss_gen', 'loss_kld'):
        # same plan, same kernels; only the order of the split-K reduce-adds / fp64 atomics may differ between runs
        assert rel(st[k], base[k]) < 2e-6, (k, st[k], base[k])
    # permuting the batch rows (with their masks and eps) permutes nothing observable: batch means
    perm = np.random.default_rng(1).permutation(64)
    st = h.grad_step(w['src'][perm], w['src'][perm], keep=w['keep'][perm], eps=w['eps'][perm])
    assert st['n_tokens'] == base['n_tokens']
    tol = 1e-5 if mode == 'fp32' else 2e-3   # bf16: rows land in other slices / MMA tile positions; fp32: reassociated sums
    for k in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(st[k], base[k]) < tol, (k, st[k], base[k])


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_valid_mode_per_sample_outputs_average_to_the_batch_terms(world, mode):
    w = world
    h = w['h'][mode]
    e = h.eval_step(w['src'], w['src'], want_pred=True)
    assert e['loss_gen_samp'].shape == (8872,) and e['loss_kld_samp'].shape == (64, 1024)
    assert np.isfinite(e['loss_gen_samp']).all() and (e['loss_gen_samp'] > 0).all()
    assert (e['loss_kld_samp'] >= -1e-6).all()                               # KL of a Gaussian is non-negative
    assert set(np.unique(e['errt_samp'])) <= {0.0, 1.0}
    assert e['pred'].min() >= 0 and e['pred'].max() < CFG['dim_tgt']
    # a train-mode step with nothing dropped and eps = 0 is the valid graph: its batch means are these means
    st = h.grad_step(w['src'], w['src'], keep=np.ones_like(w['keep']), eps=np.zeros_like(w['eps']))
    tol = 1e-5 if mode == 'fp32' else 1e-3
    assert rel(float(e['loss_gen_samp'].astype(np.float64).mean()), st['loss_gen']) < tol
    assert rel(float(e['loss_kld_samp'].astype(np.float64).mean()), st['loss_kld']) < tol
    assert abs(float(e['errt_samp'].mean()) - st['errt']) < 1e-6


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_embedding_of_a_row_is_independent_of_its_batch(world, mode):
    w = world
    h = w['h'][mode]
    mu = h.embed(w['src'])
    assert mu.shape == (64, 1024) and np.isfinite(mu).all()
    tol = 2e-5 if mode == 'fp32' else 3e-2
    scale = np.abs(mu).max()
    for i in (0, 17, 63):
        n = int((w['src'][i] != 1).sum())
        one = h.embed(w['src'][i:i + 1, :n])
        assert np.abs(one[0] - mu[i]).max() <= tol * scale, (i, np.abs(one[0] - mu[i]).max(), scale)
    sub = h.embed(w['src'][10:20])
    assert np.abs(sub - mu[10:20]).max() <= tol * scale
