"""Host-side mirror of the reference's model surface (src/model.py:48-219) over the C ABI.

The reference builds a TF-1 graph with `vAe(mode, ...)` and runs it with
`sess.run(fetches, feed_dict)`.  Here `vAe` returns the same `Record` whose fields are symbolic
`Fetch` handles; `Session.run` maps a fetch set onto the one C-ABI entry point that computes it:

    train_step                                   -> argsim_train_step   (src/train.py:118)
    errt_samp / loss_gen_samp / loss_kld_samp    -> argsim_eval_step    (src/train.py:109-110)
    z / mu with only `src` fed                   -> argsim_embed        (src/eval_embed_reason.py:38)
    state_in with `z` fed                        -> argsim_decode_init  (src/model.py:213)
    pred, state_ex with `lead`, `state_in` fed   -> argsim_decode_step  (src/model.py:215)
    step / rate_*                                -> host scalars        (src/model.py:75-80)

Like TF's variable store, all models built in one process share ONE variable set: the Session
owns weights, Adam slots, the global step and the RNG seed (src/model.py:6, src/train.py:91-96).
"""
import queue
import threading

import numpy as np

from . import _lib
from .util import Record

_MODES = ('train', 'valid', 'infer')
_state = dict(config=None, seed=0, session=None)


def set_random_seed(seed):
    """tf.set_random_seed (src/train.py:45): seeds word dropout, eps and variable init."""
    _state['seed'] = int(seed)
    if _state['session'] is not None:
        _state['session'].handle.set_seed(seed)


class Fetch:
    """a named node of the model graph; also usable as a feed key (like a tf.Tensor)."""

    def __init__(self, model, name, default=None):
        self.model, self.name, self.default = model, name, default

    def eval(self, feed_dict=None, session=None):
        sess = session or _state['session']
        if sess is None:
            raise RuntimeError('no default session; create argsim_b200.Session() first')
        return sess.run(self, feed_dict)

    def __repr__(self):
        return '<argsim_b200.Fetch %s/%s>' % (self.model.mode, self.name)

    @property
    def shape(self):
        c = self.model.config
        return {'mu': (None, c['dim_rep']), 'lv': (None, c['dim_rep']), 'z': (None, c['dim_rep'])}.get(self.name, None)


class _Prefetcher:
    """tf.data `from_generator(...).repeat(-1).prefetch(n)` (src/util_tf.py:16-23): a producer thread
    keeps `prefetch` batches ready so host tokenisation hides behind the device step."""

    def __init__(self, gen_fn, prefetch):
        self.q = queue.Queue(maxsize=max(1, prefetch))
        self.gen_fn = gen_fn
        self.err = None
        t = threading.Thread(target=self._work, daemon=True)
        t.start()

    def _work(self):
        try:
            while True:  # repeat(-1)
                n = 0
                for item in self.gen_fn():
                    self.q.put(tuple(np.ascontiguousarray(x, np.int32) for x in item))
                    n += 1
                if n == 0:
                    raise RuntimeError('batch generator yielded nothing')
        except BaseException as e:  # surfaced on the consumer side
            self.err = e
            self.q.put(None)

    def get(self):
        item = self.q.get()
        if item is None:
            raise RuntimeError('batch generator failed: %r' % (self.err,))
        return item


class _PipeSlot:
    def __init__(self, pre, index):
        self.pre, self.index = pre, index


def pipe(gen_fn, output_types=None, prefetch=1, repeat=-1, name='pipe', **kwargs):
    """see reference `util_tf.pipe`: returns one queue-backed tensor per generator output."""
    n = len(output_types) if output_types is not None else 2
    pre = _Prefetcher(gen_fn, prefetch)
    return tuple(_PipeSlot(pre, i) for i in range(n))


_FIELDS = ('src', 'tgt', 'lead', 'mu', 'lv', 'z', 'state_in', 'state_ex', 'logits', 'prob', 'pred', 'step', 'rate_keepwd',
           'rate_anneal', 'rate_update')
_LOSS_FIELDS = ('errt_samp', 'errt', 'loss_gen_samp', 'loss_gen', 'loss_kld_samp', 'loss_kld', 'loss')


def vAe(mode, src=None, tgt=None, dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, bidirectional=True,
        bidir_stacked=True, attentive=False, logit_use_embed=True, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1):
    """same signature, modes and Record fields as reference `vAe` (src/model.py:48-189)."""
    assert mode in _MODES
    config = dict(dim_tgt=dim_tgt, dim_emb=dim_emb, dim_rep=dim_rep, rnn_layers=rnn_layers, bidirectional=bidirectional,
                  bidir_stacked=bidir_stacked, attentive=attentive, logit_use_embed=logit_use_embed, accelerate=accelerate,
                  learn_rate=learn_rate, bos=bos, eos=eos)
    if _state['config'] is None:
        _state['config'] = config
    elif _state['config'] != config:
        raise ValueError('all models of one process share one variable set (tf.AUTO_REUSE, src/model.py:6); '
                         'got a different model config than the first vAe() call; use reset() to start over')
    self = Record(bos=bos, eos=eos)
    self.mode, self.config = mode, config
    for f in _FIELDS:
        setattr(self, f, Fetch(self, f))
    self.src.default, self.tgt.default = src, tgt  # placeholder_with_default (src/util_tf.py:26-37)
    if mode != 'infer':
        for f in _LOSS_FIELDS:
            setattr(self, f, Fetch(self, f))
    if mode == 'train':
        self.train_step = Fetch(self, 'train_step')
    return self


def reset():
    """forgets the model config and default session (tf.reset_default_graph analogue)."""
    if _state['session'] is not None:
        _state['session'].close()
    _state.update(config=None, session=None)


class Session:
    """tf.InteractiveSession analogue: owns the device state of the shared variable set."""

    def __init__(self, precision='bf16', device=None, nranks=None, rank=0, nccl_id=None, flags=0, config=None):
        cfg = config or _state['config']
        if cfg is None:
            raise RuntimeError('build a model with vAe(...) before creating the Session')
        if nranks is None:
            # under `python -m torch.distributed.run` every rank builds the same graph and owns one GPU: the session turns
            # data parallel by itself (batch rows sharded per step, gradients all-reduced inside the library)
            from . import parallel
            nranks, rank, local = parallel.env_world()
            if nranks > 1:
                device = local if device is None else device
                nccl_id = parallel.exchange_nccl_id(_lib.nccl_unique_id, nranks, rank)
        device = 0 if device is None else device
        prec = {'bf16': _lib.BF16, 'fp32': _lib.FP32_VALIDATE}[precision] if isinstance(precision, str) else precision
        self.handle = _lib.Handle(precision=prec, device=device, nranks=nranks, rank=rank, nccl_id=nccl_id, flags=flags, **cfg)
        self.handle.set_seed(_state['seed'])
        self.config = dict(cfg)
        self.nranks, self.rank = nranks, rank
        self._last_stats = None
        self._inflight = 0     # train steps submitted and not yet waited for (<= 2)
        # data parallel: every rank must train on rows of the SAME global batch.  'check' (default) compares a digest of
        # the ranks' batches on the first steps and then every 256th; 'broadcast' ships rank 0's batch to everybody every
        # step (what --sample needs: sentencepiece sampling is not seeded, the ranks' streams differ); None trusts the caller
        self.sync_batches = 'check'
        self._dp_steps = 0
        self.pipelined = True  # sess.run(model.train_step) returns once the step is enqueued (src/train.py:118 fetches nothing else)
        _state['session'] = self

    def _drain(self):
        while self._inflight:
            self._last_stats = self.handle.train_step_wait()
            self._inflight -= 1

    @property
    def last_stats(self):
        """statistics of the most recent training step (waits for it)."""
        self._drain()
        return self._last_stats

    def close(self):
        try:
            self._drain()
        finally:
            self.handle.close()
        if _state['session'] is self:
            _state['session'] = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------------------
    def run(self, fetches, feed_dict=None):
        single = isinstance(fetches, Fetch)
        flist = [fetches] if single else list(fetches)
        feeds = {}
        for k, v in (feed_dict or {}).items():
            if not isinstance(k, Fetch):
                raise TypeError('feed keys must be model fields, got %r' % (k,))
            feeds[k.name] = v
        model = flist[0].model
        vals = self._compute(model, [f.name for f in flist], feeds)
        out = [vals[f.name] for f in flist]
        return out[0] if single else tuple(out)

    def _feed_or_default(self, model, feeds, key):
        if key in feeds:
            return feeds[key]
        d = getattr(model, key).default
        if isinstance(d, _PipeSlot):
            return d
        return d

    def _compute(self, model, names, feeds):
        h = self.handle
        want = set(names)
        vals = {}
        scal = {'step', 'rate_keepwd', 'rate_anneal', 'rate_update'}
        if want & scal:
            step = h.step
            vals.update(_lib.schedule(step, model.config['accelerate'], model.config['learn_rate']), step=step)
            want -= scal
        if not want:
            return vals
        if 'train_step' in want:
            if want != {'train_step'}:
                raise NotImplementedError('fetch train_step on its own (src/train.py:118)')
            src = self._feed_or_default(model, feeds, 'src')
            tgt = self._feed_or_default(model, feeds, 'tgt')
            if isinstance(src, _PipeSlot) or isinstance(tgt, _PipeSlot):
                slot = src if isinstance(src, _PipeSlot) else tgt
                item = slot.pre.get()  # one dequeue feeds both placeholders (one iterator.get_next())
                if isinstance(src, _PipeSlot):
                    src = item[src.index]
                if isinstance(tgt, _PipeSlot):
                    tgt = item[tgt.index]
            if src is None or tgt is None:
                raise ValueError('train_step needs src and tgt (feed them or build the model on a pipe)')
            kw = {}
            if self.nranks > 1:
                # every rank draws the same global batch (same seed, same generator) and keeps its rows; the two loss
                # means are normalised by the GLOBAL counts (model.py:181,184), known from the lengths before launch
                from . import parallel
                if self.sync_batches == 'broadcast':
                    src, tgt = parallel.broadcast_batch(src, tgt, self.rank)
                elif self.sync_batches == 'check' and (self._dp_steps < 3 or self._dp_steps % 256 == 0):
                    parallel.assert_same_batch(src, tgt)
                self._dp_steps += 1
                src, tgt, rows, n_tok, b_glob = parallel.shard_batch(src, tgt, self.nranks, self.rank, eos=model.config.get('eos', 1))
                # rows = indices in the global batch: they key the Philox streams of word dropout and eps, so the
                # un-injected randomness of a step is the same whatever the number of ranks
                kw = dict(n_tokens_global=n_tok, b_global=b_glob, rows=rows)
            if self.pipelined:
                # submit(n+1) before wait(n): the host plan + H2D of this step overlap the device's previous step
                h.train_step_submit(src, tgt, **kw)
                self._inflight += 1
                if self._inflight == 2:
                    self._last_stats = h.train_step_wait()
                    self._inflight -= 1
            else:
                self._drain()
                self._last_stats = h.train_step(src, tgt, **kw)
            vals['train_step'] = None
            return vals
        self._drain()
        if want <= {'state_in'} and 'z' in feeds:
            vals['state_in'] = h.decode_init(feeds['z'])
            return vals
        if want <= {'pred', 'state_ex'} and 'lead' in feeds and 'state_in' in feeds:
            lead = np.asarray(feeds['lead'], np.int32)
            if lead.ndim != 2 or lead.shape[0] != 1:
                raise NotImplementedError('decode steps feed lead of shape (1, b) (src/model.py:212-215)')
            pred, state = h.decode_step(lead[0], feeds['state_in'])
            vals['pred'] = pred[None, :]   # infer mode keeps the time axis: pred is (t, b) = (1, b)
            vals['state_ex'] = state
            return vals
        src = self._feed_or_default(model, feeds, 'src')
        if isinstance(src, _PipeSlot) or src is None:
            raise ValueError('feed %s.src' % model.mode)
        if want <= {'z', 'mu'}:
            if model.mode == 'train':
                raise NotImplementedError("z of a 'train' graph is sampled; fetch it from a 'valid'/'infer' model (z = mu)")
            mu = h.embed(src)
            vals.update(z=mu, mu=mu)
            return vals
        loss_like = set(_LOSS_FIELDS) | {'pred'}
        if want <= loss_like | {'z', 'mu'}:
            if model.mode == 'infer':
                raise ValueError("loss fields do not exist in 'infer' mode (src/model.py:173)")
            if model.mode == 'train':
                raise NotImplementedError("evaluate losses on a 'valid' model; a 'train' graph applies dropout and sampling")
            tgt = self._feed_or_default(model, feeds, 'tgt')
            if tgt is None or isinstance(tgt, _PipeSlot):
                raise ValueError('feed %s.tgt' % model.mode)
            o = h.eval_step(src, tgt, want_pred='pred' in want)
            vals.update(o)
            vals['errt'] = np.float32(o['errt_samp'].mean())
            vals['loss_gen'] = np.float32(o['loss_gen_samp'].mean())
            vals['loss_kld'] = np.float32(o['loss_kld_samp'].mean())
            anneal = _lib.schedule(h.step, model.config['accelerate'], model.config['learn_rate'])['rate_anneal']
            vals['loss'] = np.float32(anneal * vals['loss_kld'] + vals['loss_gen'])
            if want & {'z', 'mu'}:
                mu = h.embed(src)
                vals.update(z=mu, mu=mu)
            return vals
        raise NotImplementedError('fetch set %r is not served by the hot-path library' % sorted(want))


def global_variables_initializer(session=None, seed=None):
    """tf.global_variables_initializer().run() (src/train.py:96): A22 initial values."""
    sess = session or _state['session']
    sess.handle.init_params(_state['seed'] if seed is None else seed)


class Saver:
    """tf.train.Saver stand-in (src/train.py:92-96,121): parameters, Adam slots and the global step.  `save` writes
    the library's own flat container, or -- `tf_format=True` -- a TF V2 bundle under the reference's variable names;
    `restore` reads either: a path with a `<path>.index` next to it is taken for a checkpoint the reference wrote
    (argsim_b200/tf_ckpt.py: container format restated without TensorFlow, name mapping unverified -- see its header)."""

    def save(self, sess, path, write_meta_graph=False, tf_format=False):
        if tf_format:
            from . import tf_ckpt
            tf_ckpt.save_reference_checkpoint(sess.handle, path, sess.config)
        else:
            sess.handle.save(path)
        return path

    def restore(self, sess, path):
        import os
        if os.path.exists(path + '.index') and not os.path.exists(path):
            from . import tf_ckpt
            self.unplaced = tf_ckpt.load_reference_checkpoint(sess.handle, path, sess.config)
        else:
            sess.handle.load(path)


def encode(sess, vae, src):
    """returns latent states: array f32 (b, dim_rep) from src: array i32 (b, t)  (src/model.py:194-201)."""
    return sess.run(vae.z, {vae.src: src})


def decode(sess, vae, z, steps=256):
    """greedy decoding of latent states (src/model.py:204-219): array i32 (b, t), t <= steps.
    As in the reference, an output that is all-eos from the first step raises (np.concatenate of [])."""
    if getattr(sess, 'handle', None) is not None and not _state.get('decode_host_loop'):
        # the whole loop on the device (argsim_decode): same tokens as the per-step path below, no host round trips
        y = sess.handle.decode(z, steps)
        if y.shape[1] == 0:
            raise ValueError('need at least one array to concatenate')   # what np.concatenate([]) raises in the reference
        return y
    x = np.full((1, len(z)), vae.bos, dtype=np.int32)
    s = sess.run(vae.state_in, {vae.z: z})
    y = []
    for _ in range(steps):
        x, s = sess.run((vae.pred, vae.state_ex), {vae.lead: x, vae.state_in: s})
        x = np.asarray(x, np.int32).reshape(1, -1)
        if np.all(x == vae.eos):
            break
        y.append(x)
    return np.concatenate(y).T
