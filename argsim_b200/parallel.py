"""Data-parallel host logic: the batch is sharded by rows across ranks,
balanced on sum and max of lengths; the two loss means (model.py:181,184) are normalised by the
GLOBAL token / row counts, which the host knows from the lengths before launch, so the only
collective on the data path is the gradient all-reduce (sum) inside the library."""
import numpy as np


def lengths(x, eos=1):
    return (np.asarray(x) != eos).sum(1)


def shard_rows(lens, nranks):
    """row indices per rank: sort by length (descending, stable) and deal round-robin in a
    boustrophedon order, so every rank gets a similar longest row and a similar token sum."""
    order = np.argsort(-np.asarray(lens), kind='stable')
    shards = [[] for _ in range(nranks)]
    for k, i in enumerate(order):
        r = k % (2 * nranks)
        r = r if r < nranks else 2 * nranks - 1 - r
        shards[r].append(int(i))
    return [np.array(sorted(s), np.int64) for s in shards]


def shard_batch(src, tgt, nranks, rank, eos=1):
    """-> (src_r, tgt_r, rows_r, n_tokens_global, b_global): this rank's rows (padding re-trimmed),
    their global row indices (RNG keying), and the global normalisers."""
    src, tgt = np.asarray(src, np.int32), np.asarray(tgt, np.int32)
    ls, lt = lengths(src, eos), lengths(tgt, eos)
    rows = shard_rows(np.maximum(ls, lt), nranks)[rank]
    s, t = src[rows], tgt[rows]
    s = np.ascontiguousarray(s[:, :max(1, int(ls[rows].max()))])
    t = np.ascontiguousarray(t[:, :max(1, int(lt[rows].max()))])
    return s, t, rows, int((lt + 1).sum()), int(len(src))
