"""-m gpu: the pipelined form of the training call (argsim_train_step_submit / argsim_train_step_wait,
include/argsim_b200.h) is the blocking argsim_train_step with the host side of step n+1 overlapped with the device's
step n: same statistics, same weights, same step counter; misuse is an error, not a hang."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(dim_tgt=256, dim_emb=64, dim_rep=128, rnn_layers=2, accelerate=1e-3, learn_rate=1e-3)


def _batches(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        b, T = int(rng.integers(3, 9)), int(rng.integers(5, 40))    # shapes change from step to step
        lens = rng.integers(1, T + 1, b)
        x = np.ones((b, T), np.int32)
        for i, l in enumerate(lens):
            x[i, :l] = rng.integers(3, CFG['dim_tgt'], l)
        out.append(x)
    return out


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_pipelined_equals_blocking(precision):
    from argsim_b200 import _lib
    prec = _lib.FP32_VALIDATE if precision == 'fp32' else _lib.BF16
    data = _batches(7, 3)
    runs = []
    for pipelined in (False, True):
        h = _lib.Handle(precision=prec, **CFG)
        h.init_params(11)
        h.set_seed(5)
        stats = []
        if not pipelined:
            for x in data:
                stats.append(h.train_step(x, x))
        else:
            h.train_step_submit(data[0], data[0])
            for x in data[1:]:
                h.train_step_submit(x, x)
                stats.append(h.train_step_wait())
            stats.append(h.train_step_wait())
        assert h.step == len(data)
        runs.append((stats, {n: h.get_param(n) for n in ('embed/embedding', 'encode/rnn1/fwd/R', 'decode/rnn/l0/W', 'latent/mu/kernel')}))
        h.close()
    (sa, pa), (sb, pb) = runs
    tol = 1e-6 if precision == 'fp32' else 2e-3     # bf16: split-K / atomic accumulation order varies from run to run
    for a, b in zip(sa, sb):
        assert a['step'] == b['step'] and a['n_tokens'] == b['n_tokens']
        for k in ('loss', 'loss_gen', 'loss_kld', 'errt', 'rate_keepwd', 'rate_anneal', 'rate_update'):
            assert abs(a[k] - b[k]) <= tol * max(1.0, abs(a[k])), (k, a, b)
    for n in pa:
        np.testing.assert_allclose(pb[n], pa[n], rtol=0, atol=1e-6 if precision == 'fp32' else 5e-4, err_msg=n)


def test_pipeline_misuse_is_an_error():
    from argsim_b200 import _lib
    h = _lib.Handle(precision=_lib.BF16, **CFG)
    h.init_params(1)
    x = _batches(1, 0)[0]
    with pytest.raises(RuntimeError):
        h.train_step_wait()                 # nothing in flight
    h.train_step_submit(x, x)
    h.train_step_submit(x, x)
    with pytest.raises(RuntimeError):
        h.train_step_submit(x, x)           # a third un-waited step
    with pytest.raises(RuntimeError):
        h.train_step(x, x)                  # the blocking call refuses to jump the queue
    mu = h.embed(x)                         # any other entry point lets the steps in flight finish first
    assert np.isfinite(mu).all()
    a, b = h.train_step_wait(), h.train_step_wait()
    assert (a['step'], b['step']) == (1, 2)
    assert h.step == 2
    h.close()
