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
    if len(src) < nranks:
        raise ValueError('data parallel: the global batch has %d rows, fewer than the %d ranks (every rank needs >= 1 row)'
                         % (len(src), nranks))
    ls, lt = lengths(src, eos), lengths(tgt, eos)
    rows = shard_rows(np.maximum(ls, lt), nranks)[rank]
    s, t = src[rows], tgt[rows]
    s = np.ascontiguousarray(s[:, :max(1, int(ls[rows].max()))])
    t = np.ascontiguousarray(t[:, :max(1, int(lt[rows].max()))])
    return s, t, rows, int((lt + 1).sum()), int(len(src))


def env_world():
    """(nranks, rank, local_rank) of a `python -m torch.distributed.run` launch; (1, 0, 0) outside one."""
    import os
    n = int(os.environ.get('WORLD_SIZE', '1'))
    if n <= 1:
        return 1, 0, 0
    return n, int(os.environ['RANK']), int(os.environ.get('LOCAL_RANK', os.environ['RANK']))


def exchange_nccl_id(make_id, nranks, rank):
    """rank 0 draws the NCCL unique id (`make_id()` -> 128 bytes), every rank returns it.  The hand-over runs over a
    torch.distributed gloo group on the launcher's MASTER_ADDR / MASTER_PORT rendezvous: host plumbing only, the
    gradients never travel this way (the library all-reduces them with NCCL over NVLink)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group('gloo', rank=rank, world_size=nranks)
    obj = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return bytes(obj[0])


def broadcast_batch(src, tgt, rank):
    """rank 0's (src, tgt) on every rank (gloo, host plumbing).  Needed whenever the batch stream is not a pure
    function of the seed: `--sample` draws segmentations from sentencepiece's unseeded per-process RNG
    (src/util_sp.py:66-111), so the ranks' own copies of a batch differ in ids AND lengths and the row sets of
    shard_rows would not partition one batch."""
    import torch
    import torch.distributed as dist
    hdr = torch.zeros(3, dtype=torch.int64)
    if rank == 0:
        src, tgt = np.ascontiguousarray(src, np.int32), np.ascontiguousarray(tgt, np.int32)
        hdr = torch.tensor([src.shape[0], src.shape[1], tgt.shape[1]], dtype=torch.int64)
    dist.broadcast(hdr, src=0)
    b, ts, tt = (int(x) for x in hdr)
    buf = torch.empty(b * (ts + tt), dtype=torch.int32)
    if rank == 0:
        buf[:b * ts] = torch.from_numpy(src.reshape(-1))
        buf[b * ts:] = torch.from_numpy(tgt.reshape(-1))
    dist.broadcast(buf, src=0)
    a = buf.numpy()
    return a[:b * ts].reshape(b, ts).copy(), a[b * ts:].reshape(b, tt).copy()


def batch_digest(src, tgt):
    """64-bit digest of a batch (shape + contents), for the cross-rank agreement check."""
    import hashlib
    h = hashlib.blake2b(digest_size=8)
    for a in (src, tgt):
        a = np.ascontiguousarray(a, np.int32)
        h.update(np.asarray(a.shape, np.int64).tobytes())
        h.update(a.tobytes())
    return int.from_bytes(h.digest(), 'little') >> 1   # fits int64


def assert_same_batch(src, tgt, what='batch'):
    """raises unless every rank holds the same batch (all-reduce MIN/MAX of a digest over gloo)."""
    import torch
    import torch.distributed as dist
    d = batch_digest(src, tgt)
    lo, hi = torch.tensor([d], dtype=torch.int64), torch.tensor([d], dtype=torch.int64)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if int(lo) != int(hi):
        raise RuntimeError('data parallel: the ranks hold different %ss (a batch stream that is not a pure function of '
                           'the seed, e.g. --sample, must be broadcast from rank 0: Session.sync_batches = "broadcast")' % what)


def row0_of(rank, nranks, b_global):
    """offset of a rank's rows in the RNG keying of word dropout / eps (distinct streams per rank)."""
    return rank * ((b_global + nranks - 1) // nranks)
