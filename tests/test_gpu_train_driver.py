"""-m gpu: the reference's user-facing flow end to end on the GPU, through the drop-in modules only
(src/train.py:40-121, src/eval_embed_reason.py:24-54, src/model.py:194-219): train a sentencepiece model
(src/util_sp.py:17-39 flags: unk=0 eos=1 bos=2), write the reference's config.json schema, run the train driver with
the reference's flags (validation summaries, checkpoint per round), resume from the checkpoint, then embed and decode
with the 'infer' graph exactly like the eval / explore scripts do."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WORDS = ('the of and to in is that it was for on are as with his they at be this from have or by one had not but what all '
         'were when we there can an your which their said if do will each about how up out them then she many some so '
         'these would other into has more her two like him see time could no make than first been its who now people my '
         'made over did down only way find use may water long little very after words called just where most know').split()


def _corpus(path, n, seed):
    rng = np.random.default_rng(seed)
    with open(path, 'w') as f:
        for _ in range(n):
            k = int(rng.integers(3, 14))
            a = rng.integers(0, len(WORDS), k)
            # a deterministic continuation makes the text learnable: word i is followed by word (3 i + 1) mod |W|
            toks = []
            for x in a:
                toks += [WORDS[x], WORDS[(3 * x + 1) % len(WORDS)]]
            f.write(' '.join(toks) + ' .\n')


def test_train_driver_resume_embed_decode(tmp_path):
    from argsim_b200 import model as M, train, util_sp
    from argsim_b200.util_np import vpack
    d = str(tmp_path)
    _corpus(d + '/train.txt', 600, 0)
    _corpus(d + '/valid.txt', 64, 1)
    util_sp.spm(d + '/vocab', d + '/train.txt', size=100)   # the toy corpus supports ~100 pieces
    vocab = util_sp.load_spm(d + '/vocab.model')
    assert (vocab.unk_id(), vocab.eos_id(), vocab.bos_id()) == (0, 1, 2)
    val = [vocab.encode_as_ids(l.strip()) for l in open(d + '/valid.txt')]
    np.save(d + '/valid.npy', vpack(val, (len(val), max(map(len, val))), vocab.eos_id(), np.int32))
    cfg = dict(paths=dict(log=d + '/log', vocab=d + '/vocab.model', train=d + '/train.txt', valid=d + '/valid.npy', ckpt=d + '/ckpt'),
               model=dict(accelerate=1e-4, learn_rate=3e-3, dim_tgt=104, dim_emb=64, dim_rep=128, rnn_layers=2, bidirectional=True,
                          bidir_stacked=True, attentive=False, logit_use_embed=True),
               train=dict(seed=0, max_len=64, batch_train=32, batch_valid=40, total_valid=64))
    json.dump(cfg, open(d + '/config.json', 'w'))
    common = ['--config', d + '/config.json', '--trial', 'unit', '--steps-per-summary', '25', '--summaries-per-round', '3',
              '--prefetch', '4']
    train.main(['--rounds', '1'] + common)
    log = [json.loads(l) for l in open(d + '/log/unit.jsonl')]
    assert [r['step'] for r in log] == [25, 50, 75]
    assert all(np.isfinite([r['step_errt'], r['step_loss_gen'], r['step_loss_kld']]).all() for r in log)
    assert log[-1]['step_loss_gen'] < log[0]['step_loss_gen'] - 0.05, log      # it learns
    assert os.path.exists(d + '/ckpt/unit0')                                   # <ckpt>/<trial><step // 10000>, train.py:121
    M._state['session'].close()
    M._state.update(config=None, session=None)

    # resume (train.py:93-94): the step counter, weights and Adam slots come back
    train.main(['--rounds', '1', '--ckpt', 'unit0'] + common)
    log = [json.loads(l) for l in open(d + '/log/unit.jsonl')]
    assert [r['step'] for r in log] == [25, 50, 75, 100, 125, 150]
    assert log[-1]['step_loss_gen'] < log[2]['step_loss_gen'] + 0.05
    sess = M._state['session']

    # eval_embed_reason.py:24-38: vAe('infer'), model.z.eval({model.src: batch}) == mu
    model = M.vAe('infer', **cfg['model'])
    valid = np.load(d + '/valid.npy')
    mu = model.z.eval({model.src: valid[:10]})
    assert mu.shape == (10, 128) and np.isfinite(mu).all()
    np.testing.assert_allclose(M.encode(sess, model, valid[:10]), mu, rtol=0, atol=0)
    # explore*.py: decode(sess, vae, z, steps) -> (b, t <= steps) token ids in range
    try:
        y = M.decode(sess, model, mu[:4], steps=12)
    except ValueError:      # every row emitted eos at once: the reference's np.concatenate([]) raises here too
        y = None
    if y is not None:
        assert y.ndim == 2 and y.shape[0] == 4 and 1 <= y.shape[1] <= 12
        assert y.min() >= 0 and y.max() < 104

    # src/eval_embed_reason.py:30-54: deterministic representation in partitions of 128 rows, and the averaged
    # representation of ONE text from its sampled segmentations (one batch of `samples` rows -> mean mu)
    from argsim_b200 import eval_embed
    texts = [l.strip() for l in open(d + '/valid.txt')][:9]
    det = eval_embed.embed_texts(sess, model, vocab, texts, batch=4)       # partitions of 4, 4, 1 rows
    assert det.shape == (9, 128)
    ids = vpack([vocab.encode_as_ids(t) for t in texts], (9, max(len(vocab.encode_as_ids(t)) for t in texts)), vocab.eos_id(), np.int32)
    np.testing.assert_allclose(det, model.z.eval({model.src: ids}), rtol=0, atol=2e-2 * np.abs(det).max())   # bf16: rows land in other MMA tiles
    avg = eval_embed.infer_avg(sess, model, vocab, texts[0], samples=32)
    assert avg.shape == (128,) and np.isfinite(avg).all()
    # sampled segmentations of a text encode the same sentence: their mean stays close to the deterministic embedding
    # compared with the spread between different texts
    d_same = np.linalg.norm(avg - det[0])
    d_other = np.median([np.linalg.norm(det[i] - det[0]) for i in range(1, 9)])
    assert d_same < d_other, (d_same, d_other)
    sess.close()
    M._state.update(config=None, session=None)

    # --sample (src/train.py:56-63): src and tgt are two independent sampled segmentations of each text
    train.main(['--rounds', '1', '--ckpt', 'unit0', '--sample'] + [('kudo' if a == 'unit' else a) for a in common])
    log = [json.loads(l) for l in open(d + '/log/kudo.jsonl')]
    assert [r['step'] for r in log] == [175, 200, 225]      # 'unit0' was overwritten by the resumed run at step 150
    assert all(np.isfinite([r['step_errt'], r['step_loss_gen'], r['step_loss_kld']]).all() for r in log)
    M._state['session'].close()
    M._state.update(config=None, session=None)


def test_sample_batches_have_independent_src_and_tgt(tmp_path):
    """the --sample feed (src/train.py:56-63, src/util_sp.py:90-111): the generator yields src != tgt of the same texts"""
    from argsim_b200 import train, util_sp
    from argsim_b200.util import Record
    d = str(tmp_path)
    _corpus(d + '/train.txt', 300, 0)
    util_sp.spm(d + '/vocab', d + '/train.txt', size=100)
    vocab = util_sp.load_spm(d + '/vocab.model')
    T = Record(dict(batch_train=16, max_len=64))
    P = Record(dict(train=d + '/train.txt'))
    gen = train.make_batch_fn(T, P, vocab, 0, True, util_sp.encode_capped, util_sp.encode_capped_sample_pair)()
    differ = 0
    for _ in range(3):
        src, tgt = next(gen)
        assert src.shape[0] == tgt.shape[0] == 16 and src.dtype == tgt.dtype == np.int32
        for a, b in zip(src, tgt):
            ta = vocab.decode_ids([int(x) for x in a if x != vocab.eos_id()])
            tb = vocab.decode_ids([int(x) for x in b if x != vocab.eos_id()])
            assert ta == tb                                  # the same text ...
            differ += int(len(a) != len(b) or (a != b).any())
    assert differ > 0                                        # ... segmented differently at least sometimes
