"""CPU, world_size 2 over gloo: the data-parallel host path.  Each rank takes its shard
(argsim_b200.parallel.shard_batch), normalises with the GLOBAL counts and all-reduces (sum) losses and
gradients -- exactly what the library does with NCCL -- and the result must equal the single-rank batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import SMALL, ragged_batch
from oracle import vae_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from argsim_b200 import parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cfg = dict(SMALL, dim_tgt=64, dim_emb=16, dim_rep=24, rnn_layers=2)
    P = O.init_params(cfg, seed=0, bias_scale=0.1)
    src = ragged_batch(10, 9, cfg['dim_tgt'], 1)
    tgt = ragged_batch(10, 8, cfg['dim_tgt'], 2)
    rng = np.random.default_rng(3)
    keep = (rng.random(tgt.shape) < 0.7).astype(np.uint8)
    eps = rng.standard_normal((10, cfg['dim_rep']))
    s, t, rows, n_glob, b_glob = parallel.shard_batch(src, tgt, world, rank)
    assert n_glob == int(((tgt != 1).sum(1) + 1).sum()) and b_glob == 10
    kr = keep[rows][:, :t.shape[1]]
    tmax = int((t != 1).sum(1).max())
    o, cache = O.forward(P, cfg, s, t, 'train', step=5000, keep=kr[:, :tmax].T, eps=eps[rows], n_tokens_global=n_glob, b_global=b_glob)
    G = O.backward(P, cfg, cache)
    flat = torch.tensor(np.concatenate([G[k].ravel() for k in sorted(G)] + [[o['loss_gen'], o['loss_kld'], o['errt']]]))
    dist.all_reduce(flat)      # the one collective of the data path: sum
    rows_all = [None] * world
    dist.all_gather_object(rows_all, rows.tolist())
    if rank == 0:
        assert sorted(sum(rows_all, [])) == list(range(10))      # shards partition the batch
        tm = int((tgt != 1).sum(1).max())
        of, cf = O.forward(P, cfg, src, tgt, 'train', step=5000, keep=keep[:, :tm].T, eps=eps)
        Gf = O.backward(P, cfg, cf)
        ref = np.concatenate([Gf[k].ravel() for k in sorted(Gf)] + [[of['loss_gen'], of['loss_kld'], of['errt']]])
        q.put(float(np.abs(flat.numpy() - ref).max() / np.abs(ref).max()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reduce_to_the_full_batch():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 1e-10, err


# ---- the session / train driver side of data parallelism, with the device library stubbed out ------------------------
class _FakeHandle:
    """stands in for argsim_b200._lib.Handle: records what the session asks the library to do."""
    made = []

    def __init__(self, **kw):
        self.kw, self.calls, self.step = kw, [], 0
        _FakeHandle.made.append(self)

    def set_seed(self, seed):
        self.seed = seed

    def train_step_submit(self, src, tgt, **kw):
        self.calls.append((np.array(src), np.array(tgt), kw))
        self.step += 1

    def train_step_wait(self):
        return dict(step=self.step)

    def close(self):
        pass


def test_session_shards_every_batch_under_torchrun(monkeypatch):
    """Session() inside a torch.distributed.run launch: device = LOCAL_RANK, the NCCL id comes from rank 0, and every
    sess.run(train_step) hands the library this rank's rows with the GLOBAL normalisers."""
    from argsim_b200 import _lib, model as M, parallel
    M.reset()
    monkeypatch.setenv('WORLD_SIZE', '2')
    monkeypatch.setenv('RANK', '1')
    monkeypatch.setenv('LOCAL_RANK', '1')
    monkeypatch.setattr(_lib, 'Handle', _FakeHandle)
    monkeypatch.setattr(parallel, 'exchange_nccl_id', lambda make, n, r: b'\x07' * 128)
    _FakeHandle.made.clear()
    cfg = dict(dim_tgt=64, dim_emb=16, dim_rep=24, rnn_layers=2)
    m = M.vAe('train', **cfg)
    sess = M.Session()
    sess.sync_batches = None      # no process group in this test; the cross-rank checks have their own test below
    h = _FakeHandle.made[-1]
    assert (h.kw['nranks'], h.kw['rank'], h.kw['device'], h.kw['nccl_id']) == (2, 1, 1, b'\x07' * 128)
    src = ragged_batch(10, 9, 64, 1)
    tgt = ragged_batch(10, 8, 64, 2)
    for _ in range(3):
        sess.run(m.train_step, {m.src: src, m.tgt: tgt})
    assert len(h.calls) == 3
    s, t, kw = h.calls[0]
    es, et, rows, n_glob, b_glob = parallel.shard_batch(src, tgt, 2, 1)
    np.testing.assert_array_equal(s, es)
    np.testing.assert_array_equal(t, et)
    assert set(kw) == {'n_tokens_global', 'b_global', 'rows'} and kw['n_tokens_global'] == n_glob and kw['b_global'] == 10
    np.testing.assert_array_equal(kw['rows'], rows)     # global row indices key the RNG streams
    other = parallel.shard_batch(src, tgt, 2, 0)[2]
    assert sorted(list(rows) + list(other)) == list(range(10))
    assert sess.last_stats == dict(step=3)      # waits for the steps in flight
    sess.close()
    M.reset()


def _id_worker(rank, world, port, q):
    from argsim_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    assert parallel.env_world() == (world, rank, rank)
    got = parallel.exchange_nccl_id(lambda: bytes(range(128)), world, rank)     # only rank 0's maker may be used
    q.put((rank, got == bytes(range(128))))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_id_hand_over_between_two_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def _sync_worker(rank, world, port, q):
    from argsim_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    same_s, same_t = ragged_batch(6, 9, 64, 1), ragged_batch(6, 7, 64, 2)
    parallel.assert_same_batch(same_s, same_t)                       # identical batches pass
    mine_s = ragged_batch(6, 9 + rank, 64, 10 + rank)                # --sample: ids AND lengths differ per rank
    mine_t = ragged_batch(6, 8, 64, 20 + rank)
    try:
        parallel.assert_same_batch(mine_s, mine_t)
        caught = False
    except RuntimeError:
        caught = True
    bs, bt = parallel.broadcast_batch(mine_s, mine_t, rank)          # rank 0's batch everywhere
    ref_s, ref_t = ragged_batch(6, 9, 64, 10), ragged_batch(6, 8, 64, 20)
    ok = caught and np.array_equal(bs, ref_s) and np.array_equal(bt, ref_t) and bs.dtype == np.int32
    parallel.assert_same_batch(bs, bt)
    rows = parallel.shard_batch(bs, bt, world, rank)[2]
    allrows = [None] * world
    dist.all_gather_object(allrows, rows.tolist())
    ok = ok and sorted(sum(allrows, [])) == list(range(6))            # the shards partition ONE batch again
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sampled_batches_are_broadcast_before_sharding():
    """with --sample every rank draws other segmentations; rank 0's batch must be THE batch."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def test_shard_batch_needs_a_row_per_rank():
    from argsim_b200 import parallel
    src = ragged_batch(3, 5, 64, 1)
    with pytest.raises(ValueError, match='fewer than'):
        parallel.shard_batch(src, src, 4, 0)
