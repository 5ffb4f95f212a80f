"""CPU: host-side mirror of the reference surface (argsim_b200/util*.py, model.py facade), the host-only
C-ABI entry points (index pipeline, schedule) against the oracle, and the library's exported symbols."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, ragged_batch
from oracle import vae_oracle as O


def test_library_exports_every_declared_symbol():
    from argsim_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'argsim_b200.h')).read()
    declared = set(re.findall(r'\b(argsim_[a-z_0-9]+)\s*\(', hdr))
    declared -= {'argsim_handle', 'argsim_config', 'argsim_step_stats'}
    assert len(declared) >= 30
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b'sm_100a' in L.argsim_version()
    # development microbenchmarks live in their own library and header, outside the product
    assert not hasattr(L, 'argsim_bench_exchange')
    dev_hdr = open(os.path.join(ROOT, 'include', 'argsim_b200_dev.h')).read()
    dev_declared = set(re.findall(r'\b(argsim_[a-z_0-9]+)\s*\(', dev_hdr))
    D = _lib.dev_lib()
    assert dev_declared == set(_lib.DEV_SIGNATURES) and all(hasattr(D, n) for n in dev_declared)


def test_no_cpu_fallback_create_fails_loudly_without_gpu(have_gpu):
    from argsim_b200 import _lib
    if have_gpu:
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CUDA device|CUDA'):
        _lib.Handle(dim_tgt=64, dim_emb=64, dim_rep=64, rnn_layers=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'argsim_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cpp', '.h', '.cuh')):
                txt = open(os.path.join(dp, f), errors='ignore').read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M), f


def test_vpack_partition_sample_match_oracle_and_reference_docstrings():
    from argsim_b200 import util_np as U
    rows = [[3, 4, 5], [6], [], [7, 8]]
    np.testing.assert_array_equal(U.vpack(rows, (4, 3), 1, np.int32), O.vpack(rows, (4, 3), 1, np.int32))
    np.testing.assert_array_equal(U.vpack(rows, (2, 4), 9, np.int32), [[3, 4, 5, 9], [6, 9, 9, 9]])
    for n, m in ((10, 3), (9, 3), (2, 5), (0, 4), (4096, 200), (128, 128)):
        for disc in (False, True):
            assert list(U.partition(n, m, disc)) == O.partition(n, m, disc), (n, m, disc)
    assert list(U.partition(10, 3)) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    a, b = U.sample(7, seed=3), O.sample(7, seed=3)
    xs = [next(a) for _ in range(21)]
    assert xs == [next(b) for _ in range(21)]
    assert sorted(xs[:7]) == list(range(7)) and xs[:7] != xs[7:14]   # epoch 2 = permutation applied twice
    rs = np.random.RandomState(3); p = list(range(7)); rs.shuffle(p)
    assert xs[:7] == p and xs[7:14] == [p[i] for i in p]


def test_record_and_comp():
    from argsim_b200.util import Record, comp, select
    r = Record(Record(a=1), b=2)
    assert r.a == 1 and r['b'] == 2 and dict(r) == {'a': 1, 'b': 2} and len(r) == 2
    assert dict(**r) == {'a': 1, 'b': 2}                 # train.py splats Records into vAe(**C)
    assert dict(select(r, 'b')) == {'b': 2}
    assert comp(lambda x: x + 1, lambda x: x * 2)(5) == 11
    assert comp(np.mean, np.concatenate)([np.ones(2), np.zeros(2)]) == 0.5


def test_schedule_entry_point_matches_oracle_fp32():
    from argsim_b200 import _lib
    for step in (0, 1, 250, 10000, 123456, 4000000):
        s, o = _lib.schedule(step), O.schedule(step)
        for k in ('rate_keepwd', 'rate_anneal', 'rate_update'):
            assert abs(float(s[k]) - float(o[k])) <= 2e-7 * max(1.0, abs(float(o[k]))), (step, k)


@pytest.mark.parametrize('seed', range(12))
def test_plan_batch_bit_exact_vs_oracle(seed):
    """trim / lead / gold / mask / boolean_mask order / final-state index: the C++ plan vs the numpy oracle"""
    from argsim_b200 import _lib
    rng = np.random.default_rng(seed)
    b = int(rng.integers(1, 40))
    src = ragged_batch(b, int(rng.integers(1, 30)), 500, seed)
    tgt = ragged_batch(b, int(rng.integers(1, 30)), 500, seed + 100)
    if seed % 3 == 0:   # extra all-eos columns on the right must be trimmed away
        src = np.concatenate([src, np.ones((b, 4), np.int32)], 1)
    keep = (rng.random(tgt.shape) < 0.6).astype(np.uint8) if seed % 2 else None
    p = _lib.plan_batch(src, tgt, bos=2, eos=1, keep=keep)
    st, ms, ls = O.trim(np.ascontiguousarray(src.T), 1)
    tt, mt, lt = O.trim(np.ascontiguousarray(tgt.T), 1)
    np.testing.assert_array_equal(p['len_src'], ls)
    np.testing.assert_array_equal(p['len_tgt'], lt)
    assert p['S'] == ls.sum() and p['N'] == (lt + 1).sum() and p['Tmax_src'] == ls.max() and p['Tmax_dec'] == lt.max() + 1
    lead, gold, msk = O.decoder_io(tt, mt, 2, 1, None if keep is None else keep[:, :tt.shape[0]].T)
    # packed order -> reference (boolean_mask) order through ref_row
    ref_lead, ref_gold = np.empty(p['N'], np.int32), np.empty(p['N'], np.int32)
    ref_lead[p['ref_row']] = p['lead']
    ref_gold[p['ref_row']] = p['gold']
    np.testing.assert_array_equal(ref_lead, lead[msk])
    np.testing.assert_array_equal(ref_gold, gold[msk])
    assert sorted(p['ref_row'].tolist()) == list(range(p['N']))
    # packed encoder ids: step t holds the rows with len > t in sorted (descending, stable) order
    order = np.argsort(-ls, kind='stable')
    np.testing.assert_array_equal(p['perm_src'], order)
    ids = np.concatenate([st[t, order[ls[order] > t]] for t in range(st.shape[0])])
    np.testing.assert_array_equal(p['ids_src'], ids)
    off = np.concatenate([[0], np.cumsum([(ls > t).sum() for t in range(st.shape[0])])])
    inv = np.argsort(order)
    np.testing.assert_array_equal(p['enc_last'], off[ls - 1] + inv)


def test_plan_batch_rejects_contract_violations():
    from argsim_b200 import _lib
    ok = np.array([[4, 5, 1]], np.int32)
    with pytest.raises(ValueError, match='no non-eos'):
        _lib.plan_batch(np.array([[1, 1, 1]], np.int32), ok)
    with pytest.raises(ValueError, match='trim'):
        _lib.plan_batch(np.array([[4, 1, 5]], np.int32), ok)
    with pytest.raises(ValueError, match='trim'):
        _lib.plan_batch(ok, np.array([[1, 4, 5]], np.int32))
    p = _lib.plan_batch(ok, np.array([[1, 1, 1]], np.int32))      # empty target is legal: one row (bos -> eos)
    assert p['N'] == 1 and p['lead'].tolist() == [2] and p['gold'].tolist() == [1]


def test_philox_numpy_restatement_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (Salmon et al. 2011, kat_vectors)"""
    from argsim_b200 import rng
    z = np.uint32(0)
    out = rng.philox4x32_10(z, z, z, z, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.uint32(0xffffffff)
    out = rng.philox4x32_10(f, f, f, f, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = rng.philox4x32_10(np.uint32(0x243f6a88), np.uint32(0x85a308d3), np.uint32(0x13198a2e), np.uint32(0x03707344),
                            0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    m = rng.keep_mask(50, 40, 0.7, seed=1, step=2)
    assert m.shape == (50, 40) and 0.6 < m.mean() < 0.8


def test_synth_workloads_match_survey_8d():
    from argsim_b200.synth import synth_batch
    x = synth_batch(64, 'iac', 8192, seed=0)
    assert x.shape == (64, 512) and int((x != 1).sum()) == 8808        # C1: S = 8,808, N = S + 64 = 8,872
    np.testing.assert_array_equal(x, O.synth_batch(64, 'iac', 8192, seed=0))
    x = synth_batch(512, 'iac', 8192, seed=0)
    assert int((x != 1).sum()) == 66266                                  # C3 (SURVEY quotes 66,780 from its own draw)
    x = synth_batch(4096, 'ibm', 8192, seed=0)
    assert int((x != 1).sum()) == 138114 and x.shape[1] == 153         # C4 (SURVEY: 138,726)
    assert x.min() >= 1 and not np.isin(x, (0, 2)).any()


def test_facade_surface_and_fetch_routing():
    """vAe returns the reference's Record fields (model.py:73-189); Session.run routes fetch sets."""
    import argsim_b200.model as M
    M.reset()
    C = dict(accelerate=1e-4, learn_rate=1e-3, dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3,
             bidirectional=True, bidir_stacked=True, attentive=False, logit_use_embed=True)   # config.json "model"
    mv = M.vAe('valid', **C)
    mt = M.vAe('train', src=None, tgt=None, **C)
    mi = M.vAe('infer', **C)
    ref_fields = ['bos', 'eos', 'step', 'rate_keepwd', 'rate_anneal', 'rate_update', 'src', 'tgt', 'lead', 'mu', 'lv', 'z',
                  'state_in', 'state_ex', 'logits', 'prob', 'pred']
    loss_fields = ['errt_samp', 'errt', 'loss_gen_samp', 'loss_gen', 'loss_kld_samp', 'loss_kld', 'loss']
    for f in ref_fields:
        assert f in mv and f in mt and f in mi, f
    for f in loss_fields:
        assert f in mv and f in mt and f not in mi, f
    assert 'train_step' in mt and 'train_step' not in mv
    assert mv.bos == 2 and mv.eos == 1 and mv.mu.shape[-1] == 1024
    with pytest.raises(ValueError):
        M.vAe('valid', **dict(C, attentive=True))     # a different graph (src/model.py:136-145): its own variable set
    with pytest.raises(ValueError):
        M.vAe('valid', **dict(C, dim_rep=512))        # one variable set per process
    with pytest.raises(AssertionError):
        M.vAe('test', **C)
    M.reset()
    assert M.vAe('valid', **dict(C, attentive=True)).config['attentive'] is True
    M.reset()


def test_pipe_prefetcher_repeats_and_feeds_pairs():
    import argsim_b200.model as M

    def gen():
        for i in range(3):
            a = np.full((2, 3), i, np.int32)
            yield a, a + 10
    src, tgt = M.pipe(gen, (np.int32, np.int32), prefetch=2)
    got = [src.pre.get() for _ in range(7)]        # repeat(-1): wraps around after 3
    assert [int(g[0][0, 0]) for g in got] == [0, 1, 2, 0, 1, 2, 0]
    assert int(got[0][tgt.index][0, 0]) == 10


def test_shard_rows_balanced_and_complete():
    from argsim_b200 import parallel
    from argsim_b200.synth import synth_batch
    x = synth_batch(512, 'iac', 8192, seed=0)
    lens = parallel.lengths(x)
    for n in (1, 2, 4, 8):
        sh = parallel.shard_rows(lens, n)
        assert sorted(np.concatenate(sh).tolist()) == list(range(512))
        tok = [int(lens[s].sum()) for s in sh]
        assert max(tok) - min(tok) <= 0.02 * sum(tok) / n + 512, tok
        assert all(len(s) == 512 // n for s in sh)
        assert min(int(lens[s].max()) for s in sh) >= 0.9 * lens.max()
    s, t, rows, ntok, bglob = parallel.shard_batch(x, x, 4, 1)
    assert ntok == int((lens + 1).sum()) == 66266 + 512 and bglob == 512 and s.shape[0] == 128


def test_eval_embed_flows_mirror_the_reference_scripts():
    """src/eval_embed_reason.py:30-54 on a stub session: partitions of `batch` rows, eos-padded int32 matrices,
    and the mean over sampled segmentations of one text."""
    from argsim_b200 import eval_embed

    class Vocab:
        def eos_id(self):
            return 1

        def encode_as_ids(self, text):
            return [3 + len(w) for w in text.split()]

        def sample_encode_as_ids(self, text, nbest, alpha):
            assert (nbest, alpha) == (-1, 0.5)            # src/util_sp.py:66-87
            self.n = getattr(self, 'n', 0) + 1
            ids = self.encode_as_ids(text)
            return ids + [9] * (self.n % 3)               # segmentations of different lengths

    class Model:
        z, src = 'z', 'src'

    class Sess:
        def __init__(self):
            self.feeds = []

        def run(self, fetch, feed):
            assert fetch == 'z'
            x = feed['src']
            assert x.dtype == np.int32 and x.ndim == 2
            self.feeds.append(x)
            return np.stack([(x != 1).sum(1), x.sum(1)], 1).astype(np.float32)

    texts = ['a bb ccc', 'dd', 'e f g h i', 'jj kk', 'l']
    sess = Sess()
    out = eval_embed.embed_texts(sess, Model, Vocab(), texts, batch=2)
    assert [f.shape for f in sess.feeds] == [(2, 5), (2, 5), (1, 5)]       # one vpack for all texts, partitions of 2
    assert out.shape == (5, 2) and list(out[:, 0]) == [3, 1, 5, 2, 1]
    assert (sess.feeds[0][1] == [5, 1, 1, 1, 1]).all()                      # eos padding
    sess = Sess()
    avg = eval_embed.infer_avg(sess, Model, Vocab(), 'a bb ccc', samples=6)
    assert len(sess.feeds) == 1 and sess.feeds[0].shape == (6, 5)           # ONE batch of `samples` rows
    assert avg.shape == (2,) and abs(avg[0] - np.mean([3 + (k % 3) for k in range(1, 7)])) < 1e-6
    assert eval_embed.embed_texts_sampled(Sess(), Model, Vocab(), texts[:2], samples=4).shape == (2, 2)


def test_keep_mask_streams_are_keyed_by_global_row():
    from argsim_b200 import rng
    a = rng.keep_mask(6, 11, 0.6, 1234, 77, row0=0)
    b = rng.keep_mask(3, 11, 0.6, 1234, 77, rows=[4, 0, 5])
    np.testing.assert_array_equal(b, a[[4, 0, 5]])                          # a row's stream follows its global index
    np.testing.assert_array_equal(rng.keep_mask(3, 11, 0.6, 1234, 77, row0=3), a[3:6])


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs next to the GPU arm): one JSON line with the same metric, unit
    and config.workload as the repo arm, impl = reference, a cpu_baseline block and an e2e block without copies.  A time
    budget truncates the steps here (labelled as such in the line); the driver's run is full length."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, ARGSIM_REF_BUDGET_S='3', OMP_NUM_THREADS='4')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    sys.path.insert(0, ROOT)
    import bench
    assert line['impl'] == 'reference' and line['metric'] == bench.METRIC and line['unit'] == 'sequences/s'
    assert line['higher_is_better'] is True and line['vs_baseline'] is None and line['n_gpus'] == 1
    assert line['config']['workload'] == bench.workload_config(1)['workload']
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e']['value'] == line['value'] and line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['value'] > 0 and 'EXTRAPOLATED' in line['cpu_baseline']['sample']
