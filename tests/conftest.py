import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


SMALL = dict(dim_tgt=256, dim_emb=64, dim_rep=128, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)


def ragged_batch(b, tmax, vocab, seed, eos=1, tmin=1):
    """eos-padded int32 (b, max len) batch with ragged lengths in [tmin, tmax], ids in [3, vocab)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(tmin, tmax + 1, b)
    lens[rng.integers(0, b)] = tmax
    out = np.full((b, int(lens.max())), eos, np.int32)
    for i, n in enumerate(lens):
        out[i, :n] = rng.integers(3, vocab, n)
    return out


@pytest.fixture(scope='session')
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
