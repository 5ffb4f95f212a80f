"""numpy batch utilities of the data feed, same names and results as reference src/util_np.py:5-33
(the bit-exact index pipeline of SURVEY.md row 6).  `batch_sample` / `batch` of the reference are
unused by the hot path and not provided."""
import numpy as np


def vpack(arrays, shape, fill, dtype=None):
    """stacks 1-d `arrays` of different lengths into a (rows, width) matrix, padding each row
    at the end with `fill`; extra arrays beyond `shape[0]` are ignored (src/util_np.py:5-13)."""
    out = np.full(shape, fill, dtype)
    for row, arr in zip(out, arrays):
        row[:len(arr)] = arr
    return out


def partition(n, m, discard=False):
    """yields (i, j) index pairs cutting range(n) into consecutive pieces of `m`; the last,
    shorter piece is yielded too unless `discard` (src/util_np.py:16-24)."""
    full = n // m
    for k in range(full):
        yield k * m, (k + 1) * m
    if n % m and not discard:
        yield full * m, n


def sample(n, seed=0):
    """endless stream of indices in [0, n).  Each epoch reseeds numpy's global RNG with `seed`
    and shuffles the SAME list in place, so epoch k is the k-fold composition of one permutation
    (src/util_np.py:27-33) -- reproduced on purpose, the batch order depends on it."""
    order = list(range(n))
    while True:
        np.random.seed(seed)
        np.random.shuffle(order)
        yield from order
