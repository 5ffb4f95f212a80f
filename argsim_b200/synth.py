"""Synthetic token batches shaped like the reference's corpora: there is no
network for datasets, so benchmarks and parity runs draw lengths from distributions fitted to the
reference's own published posts (docs/results_iac/clustering.csv) and Zipf-distributed ids."""
import numpy as np


def synth_batch(b, kind='iac', vocab=8192, seed=0, eos=1, cap=None):
    """int32 (b, max len) eos-padded batch.
    'iac' : lengths clip(rint(lognormal(ln 87, 1.0)), 1, 512)   (median 87, p90 334, 4.4 % at cap)
    'ibm' : lengths clip(rint(lognormal(ln 30, 0.5)), 1, 256)   (sentence-level arguments)
    'full': every length == cap                                    (scaled stress config)
    ids: Zipf(1.0) over 3..vocab-1 ('full': uniform); 0=unk 1=eos 2=bos never occur inside a row."""
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
    out = np.full((b, int(lens.max())), eos, np.int32)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out
