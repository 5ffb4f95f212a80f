"""-m gpu: the tensor-memory recurrence (csrc/gru_tc.cu: tcgen05.mma in TS form, R_own resident in TMEM, h staged in
swizzled shared memory from the LL exchange) -- first its building blocks against numpy, then the whole kernel against
the per-step generic GRU on the same bf16 GEMMs, for every rows-per-MMA variant (N = 16, 32, 64, 128)."""
import numpy as np
import pytest

from conftest import ragged_batch
from test_gpu_parity import _mk, _inject, rel

pytestmark = pytest.mark.gpu


def _bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize('N,K,nacc', [(16, 64, 1), (16, 512, 1), (16, 512, 8), (32, 128, 2), (32, 512, 8), (64, 256, 4), (128, 512, 2)])
def test_ts_form_mma_matches_numpy(N, K, nacc):
    from argsim_b200 import _lib
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    D = _lib.test_ts_mma(A, B, nacc=nacc)
    ref = _bf16(A).astype(np.float64) @ _bf16(B).astype(np.float64).T
    err = np.abs(D - ref).max() / np.abs(ref).max()
    assert err < 1e-5, (N, K, err)
    # a permutation-sensitive check: unit vectors pick single elements (catches lane / column / swizzle mix-ups that a
    # random-matrix norm test could hide behind a lucky symmetry)
    A2 = np.zeros((128, K), np.float32)
    A2[np.arange(128), (np.arange(128) * 7) % K] = 1.0
    B2 = (np.arange(N * K, dtype=np.float32).reshape(N, K) % 251) / 8.0       # exactly representable in bf16
    D2 = _lib.test_ts_mma(A2, B2, nacc=nacc)
    np.testing.assert_array_equal(D2, B2[:, (np.arange(128) * 7) % K].T)


# mode 1: forward recurrences on gru_tc.cu, backward on gru_mma.cu; 3: both on gru_tc.cu; 4 (the default): whole-layer
# forward launches with many rows per slice take the TMA-fed tensor-memory kernel (flags + hs as the exchange), the rest mma.sync
@pytest.mark.parametrize('mode', [1, 3, 4, 12])
@pytest.mark.parametrize('b,tmax', [(3, 9), (64, 40), (100, 23), (150, 23), (257, 12), (512, 9), (500, 33)])
def test_tensor_memory_recurrence_matches_generic(monkeypatch, b, tmax, mode):
    """recurrences on gru_tc.cu (ARGSIM_GRU_TC) against the per-step generic GRU: losses, every gradient (the gate cache
    and hs the backward reads come from the new forward kernel), mu"""
    from argsim_b200 import _lib
    cfg = dict(dim_tgt=1024, dim_emb=512, dim_rep=256, rnn_layers=2, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
    hg, P = _mk(cfg, _lib.BF16, flags=4)
    if mode == 12:   # mode 1 with the warp-specialised LL forward kernel (k_gru_tc_fwd2) instead of the block-synchronous one
        monkeypatch.setenv('ARGSIM_GRU_TC_FWD', '2')
        mode = 1
    monkeypatch.setenv('ARGSIM_GRU_TC', str(mode))
    hm, _ = _mk(cfg, _lib.BF16, flags=0)
    src = ragged_batch(b, tmax, cfg['dim_tgt'], 60 + b)
    tgt = ragged_batch(b, max(2, tmax - 3), cfg['dim_tgt'], 61 + b)
    keep, eps = _inject(cfg, tgt, 62)
    for h in (hg, hm):
        h.step = 15000
    np.testing.assert_allclose(hm.embed(src), hg.embed(src), rtol=0, atol=2e-2)
    a = hg.grad_step(src, tgt, keep=keep, eps=eps)
    m = hm.grad_step(src, tgt, keep=keep, eps=eps)
    for name in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(m[name], a[name]) < 2e-3, (name, m[name], a[name])
    bad = {}
    for k in P:
        g, r = hm.get_grad(k).astype(np.float64).ravel(), hg.get_grad(k).astype(np.float64).ravel()
        if np.linalg.norm(r) < 1e-12:
            assert np.linalg.norm(g) < 1e-9, k
            continue
        cos = g @ r / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
        ratio = np.linalg.norm(g) / np.linalg.norm(r)
        if not (cos > 0.998 and abs(ratio - 1) < 0.02):
            bad[k] = (round(float(cos), 4), round(float(ratio), 4))
    assert not bad, bad
    # a second step on the same handles: the exchange buffers' tags / parities carry over between launches
    a = hg.grad_step(src, tgt, keep=keep, eps=eps)
    m = hm.grad_step(src, tgt, keep=keep, eps=eps)
    assert rel(m['loss'], a['loss']) < 2e-3


@pytest.mark.parametrize('mode', [1, 3])
@pytest.mark.parametrize('seg,b,tmax', [(8, 40, 37), (5, 130, 21)])
def test_tensor_memory_recurrence_in_the_decoder_wavefront(monkeypatch, seg, b, tmax, mode):
    """time-segmented launches (state hand-over through hT / h0, one stream per decoder layer) on the tcgen05 kernel"""
    from argsim_b200 import _lib
    cfg = dict(dim_tgt=1024, dim_emb=512, dim_rep=256, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
    hg, P = _mk(cfg, _lib.BF16, flags=4)
    monkeypatch.setenv('ARGSIM_DEC_SEG', str(seg))
    monkeypatch.setenv('ARGSIM_ENC_SEG', str(2 * seg))      # the encoder's BPTT as two chains of segment launches too
    monkeypatch.setenv('ARGSIM_GRU_TC', str(mode))
    hm, _ = _mk(cfg, _lib.BF16, flags=0)
    src = ragged_batch(b, tmax, cfg['dim_tgt'], 70 + b)
    tgt = ragged_batch(b, tmax, cfg['dim_tgt'], 71 + b)
    keep, eps = _inject(cfg, tgt, 72)
    for h in (hg, hm):
        h.step = 15000
    a = hg.grad_step(src, tgt, keep=keep, eps=eps)
    m = hm.grad_step(src, tgt, keep=keep, eps=eps)
    for name in ('loss', 'loss_gen', 'loss_kld'):
        assert rel(m[name], a[name]) < 2e-3, (name, m[name], a[name])
    bad = {}
    for k in P:
        g, r = hm.get_grad(k).astype(np.float64).ravel(), hg.get_grad(k).astype(np.float64).ravel()
        if np.linalg.norm(r) < 1e-12:
            continue
        cos = g @ r / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
        ratio = np.linalg.norm(g) / np.linalg.norm(r)
        if not (cos > 0.998 and abs(ratio - 1) < 0.02):
            bad[k] = (round(float(cos), 4), round(float(ratio), 4))
    assert not bad, bad
    e1, e2 = hm.eval_step(src, tgt), hg.eval_step(src, tgt)
    assert rel(e1['loss_gen_samp'].mean(), e2['loss_gen_samp'].mean()) < 2e-3
