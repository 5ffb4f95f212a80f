"""-m gpu: the tcgen05/TMA GEMM against numpy on bf16-rounded operands, all four operand layouts
(K-major / MN-major), ragged shapes (TMA zero fill), bias, alpha, accumulate and split-K."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bf16_round(x):
    import torch
    return torch.tensor(np.asarray(x, np.float32)).bfloat16().float().numpy()


def _case(M, N, K, a_mn, b_mn, bias=False, alpha=1.0, acc=False, seed=0):
    from argsim_b200 import _lib
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    bi = rng.standard_normal(N).astype(np.float32) if bias else None
    C0 = rng.standard_normal((M, N)).astype(np.float32) if acc else None
    ref = alpha * (bf16_round(A).astype(np.float64) @ bf16_round(B).astype(np.float64).T)
    if bias:
        ref = ref + bi
    if acc:
        ref = ref + C0
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    out, ms = _lib.test_gemm(1, Ain, Bin, a_mn, b_mn, bias=bi, alpha=alpha, C0=C0)
    err = np.abs(out - ref).max() / (np.abs(ref).max() + 1e-9)
    assert err < 2e-3, (M, N, K, a_mn, b_mn, err)
    if not acc:   # bf16-typed output (activations): same product, rounded once to bf16
        outh, _ = _lib.test_gemm(2, Ain, Bin, a_mn, b_mn, bias=bi, alpha=alpha)
        errh = np.abs(outh - ref).max() / (np.abs(ref).max() + 1e-9)
        assert errh < 6e-3, (M, N, K, a_mn, b_mn, errh)
    sim, _ = _lib.test_gemm(0, Ain, Bin, a_mn, b_mn, bias=bi, alpha=alpha, C0=C0)   # SIMT fp32 twin
    ref32 = alpha * (A.astype(np.float64) @ B.astype(np.float64).T) + (bi if bias else 0) + (C0 if acc else 0)
    assert np.abs(sim - ref32).max() / (np.abs(ref32).max() + 1e-9) < 1e-5
    return ms


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_layouts_one_tile(a_mn, b_mn):
    _case(128, 128, 64, a_mn, b_mn)
    _case(128, 128, 256, a_mn, b_mn, seed=1)


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_layouts_ragged(a_mn, b_mn):
    _case(200, 136, 72, a_mn, b_mn, bias=True, alpha=0.5)
    _case(64, 8, 40, a_mn, b_mn, seed=2)
    _case(8, 264, 520, a_mn, b_mn, acc=True, seed=3)   # MN-major operands need ld % 8 == 0 (TMA 16-byte strides)


@pytest.mark.parametrize('shape', [
    (8192, 512, 171, 1, 1, dict(acc=True, alpha=512 ** -0.5)),   # tied-embedding wgrad, K = ragged row count
    (171, 512, 8192, 0, 1, dict(alpha=512 ** -0.5)),             # dho = dlogits . E
    (171, 512, 1536, 0, 1, {}),                                  # GRU dgrad, split-K into a fresh C
    (1536, 512, 171, 1, 1, dict(acc=True)),                      # GRU wgrad small K
    (171, 1536, 512, 0, 0, dict(bias=True)),                     # GRU input projection
    (104, 1024, 3072, 0, 1, {}),                                 # encoder dgrad through [W_f;W_b]
    (3072, 1024, 104, 1, 1, dict(acc=True)),                     # encoder wgrad
    (16, 512, 1024, 0, 0, {}),                                   # dz = dhx . Kex^T
])
def test_step_shapes(shape):
    M, N, K, a_mn, b_mn, kw = shape
    _case(M, N, K, a_mn, b_mn, **kw)


def test_model_shapes():
    _case(1000, 1536, 512, 0, 0, bias=True)            # GRU input projection
    _case(600, 2048, 512, 0, 0, alpha=512 ** -0.5)     # vocab projection slice
    _case(1536, 512, 3000, 1, 1, acc=True)             # wgrad, split-K + accumulate
    _case(1536, 512, 3001 - 1, 1, 1)                   # wgrad, split-K into a fresh C
    _case(900, 512, 3072, 0, 1)                        # dgrad through [W_f;W_b]
    _case(64, 1024, 1024, 0, 1, bias=True)             # latent affine, tiny M
    _case(1024, 512, 64, 1, 1)                         # latent wgrad, K = batch


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_persistent_many_tiles(a_mn, b_mn):
    # more 128 x 256 work units than SMs: every CTA walks several tiles through both TMEM accumulator buffers,
    # ragged in M, N and K (TMA zero fill on loads, clipping on stores)
    _case(2504, 4104, 200, a_mn, b_mn, bias=True, alpha=0.25, seed=5)   # MN-major operands need ld % 8 == 0
    _case(3000, 2048, 72, a_mn, b_mn, acc=True, seed=6)


def test_split_k_units():
    _case(304, 520, 4096, 1, 1, acc=True, seed=7)     # few tiles, many k-blocks: split-K through TMA reduce-add
    _case(256, 256, 8192, 0, 0, seed=8)
