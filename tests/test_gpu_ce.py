"""-m gpu: the fused softmax cross-entropy kernels (src/model.py:170-181 and the gradient tf.gradients derives,
SURVEY A12-A14, A18) against numpy fp64 on the same (bf16-rounded) logits: loss per row, argmax with tf.argmax's
lowest-index tie rule, error flag, fp64 sums, in-place gradient (softmax - onehot) * gscale.  Covers the
register-resident bf16 path (V = 2048 * {1,2,4,8,16}), the shared-memory bf16 path (any V) and the fp32
validation kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bf16_round(x):
    import torch
    return torch.tensor(np.asarray(x, np.float32)).bfloat16().float().numpy()


def _ref(x, lab, gscale):
    x = x.astype(np.float64)
    mx = x.max(1, keepdims=True)
    e = np.exp(x - mx)
    s = e.sum(1, keepdims=True)
    loss = (np.log(s) + mx)[:, 0] - x[np.arange(len(x)), lab]
    pred = x.argmax(1)                      # numpy: first occurrence = lowest index, like tf.argmax
    grad = e / s
    grad[np.arange(len(x)), lab] -= 1.0
    return loss, pred, grad * gscale


def _check(n, V, bf16, seed=0, scale=3.0, gscale=None):
    from argsim_b200 import _lib
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, V)) * scale).astype(np.float32)
    lab = rng.integers(0, V, n).astype(np.int32)
    if n >= 8:
        # ties for the maximum: the lowest index must win; label at the first / last column; label == argmax
        x[0, [5, V - 1]] = 40.0
        x[1, [V - 1, V // 2 + 3, V // 2 + 2]] = 37.0
        x[2, :] = 1.5                       # everything ties -> index 0
        lab[3], lab[4] = 0, V - 1
        x[5, lab[5]] = 50.0                 # confident and right: gradient at the label ~ -0
        x[6, (lab[6] + 1) % V] = 50.0       # confident and wrong
        x[7, :] = -0.0
        x[7, min(9, V - 1)] = 0.0           # +0 == -0: tie, index 0 wins
    if bf16:
        x = bf16_round(x)
    gscale = 1.0 / n if gscale is None else gscale
    out = _lib.test_softmax_ce(x, lab, gscale=gscale, bf16=bf16)
    loss, pred, grad = _ref(x, lab, gscale)
    np.testing.assert_array_equal(out['pred'], pred)                      # index work: bit exact
    np.testing.assert_array_equal(out['err_samp'], (pred != lab).astype(np.float32))
    tol = 2e-5 if not bf16 else 1e-4                                      # loss is fp32 arithmetic in both modes
    np.testing.assert_allclose(out['loss_samp'], loss, rtol=tol, atol=tol)
    np.testing.assert_allclose(out['stats'][0], loss.sum(), rtol=1e-5)
    assert out['stats'][1] == (pred != lab).sum()
    if bf16:
        # gradient is stored as bf16 (exp kept as bf16 between the passes: <= 1 ulp = 2^-7 relative), label element fp32->bf16
        err = np.abs(out['grad'] - grad)
        err[np.arange(n), lab] = 0.0        # label elements: checked below (softmax - 1 cancels when the model is right)
        assert (err <= np.abs(grad) * 2.0 ** -7 + 1e-30).all(), (err / (np.abs(grad) + 1e-30)).max()
        assert (np.abs(out['grad'][np.arange(n), lab] - grad[np.arange(n), lab])
                <= np.abs(grad[np.arange(n), lab]) * 2.0 ** -8 + gscale * 2e-6).all()
    else:
        np.testing.assert_allclose(out['grad'], grad, rtol=2e-5, atol=1e-7 * gscale)
    # argmax-only / no-gradient mode (valid + infer graphs): logits untouched
    out2 = _lib.test_softmax_ce(x, lab, gscale=gscale, bf16=bf16, write_grad=False)
    np.testing.assert_array_equal(out2['pred'], pred)
    np.testing.assert_array_equal(out2['grad'], x)
    out3 = _lib.test_softmax_ce(x, None, gscale=gscale, bf16=bf16, write_grad=False)
    np.testing.assert_array_equal(out3['pred'], pred)


@pytest.mark.parametrize('V', [2048, 4096, 8192, 16384, 32768])
def test_ce_bf16_register_path(V):
    _check(64, V, True, seed=V)


@pytest.mark.parametrize('V', [8, 1000, 8191, 8200])
def test_ce_bf16_smem_path(V):
    _check(33, V, True, seed=V)


@pytest.mark.parametrize('V', [512, 8192, 5001])
def test_ce_fp32(V):
    _check(40, V, False, seed=V)


def test_ce_one_row_and_many_rows():
    _check(1, 8192, True, seed=1)
    _check(3000, 8192, True, seed=2, gscale=0.37)
