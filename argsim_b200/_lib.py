"""ctypes binding of libargsim_b200.so (include/argsim_b200.h).  Thin: plain pointers and sizes,
errors become Python exceptions.  The library is built in-tree by argsim_b200/build.py; there
is no CPU fallback -- `Handle(...)` raises when no CUDA device is present."""
import ctypes as C
import os

import numpy as np

from . import build as _build

FP32_VALIDATE, BF16 = 0, 1
FLAG_GENERIC_GRU, FLAG_KERNEL_TIMERS = 4, 8


class Config(C.Structure):
    _fields_ = [('dim_tgt', C.c_int32), ('dim_emb', C.c_int32), ('dim_rep', C.c_int32), ('rnn_layers', C.c_int32),
                ('bidirectional', C.c_int32), ('bidir_stacked', C.c_int32), ('attentive', C.c_int32),
                ('logit_use_embed', C.c_int32), ('accelerate', C.c_float), ('learn_rate', C.c_float),
                ('bos', C.c_int32), ('eos', C.c_int32), ('precision', C.c_int32), ('max_batch', C.c_int32),
                ('max_len', C.c_int32), ('device', C.c_int32), ('nranks', C.c_int32), ('rank', C.c_int32),
                ('nccl_id', C.c_uint8 * 128), ('flags', C.c_int32)]


class StepStats(C.Structure):
    _fields_ = [('loss', C.c_float), ('loss_gen', C.c_float), ('loss_kld', C.c_float), ('errt', C.c_float),
                ('rate_keepwd', C.c_float), ('rate_anneal', C.c_float), ('rate_update', C.c_float),
                ('n_tokens', C.c_int64), ('step', C.c_int64)]


_lib = None
_i32p, _f32p, _u8p, _i64p = (C.POINTER(t) for t in (C.c_int32, C.c_float, C.c_uint8, C.c_int64))

# name -> (restype, argtypes); every symbol include/argsim_b200.h declares
SIGNATURES = {
    'argsim_version': (C.c_char_p, []),
    'argsim_last_error': (C.c_char_p, [C.c_void_p]),
    'argsim_nccl_unique_id': (C.c_int, [_u8p]),
    'argsim_create': (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    'argsim_destroy': (None, [C.c_void_p]),
    'argsim_init_params': (C.c_int, [C.c_void_p, C.c_uint64]),
    'argsim_param_count': (C.c_int, [C.c_void_p, _i32p]),
    'argsim_param_info': (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), _i32p, _i64p]),
    'argsim_get_param': (C.c_int, [C.c_void_p, C.c_char_p, _f32p]),
    'argsim_set_param': (C.c_int, [C.c_void_p, C.c_char_p, _f32p]),
    'argsim_get_grad': (C.c_int, [C.c_void_p, C.c_char_p, _f32p]),
    'argsim_get_opt_state': (C.c_int, [C.c_void_p, C.c_char_p, _f32p, _f32p]),
    'argsim_set_opt_state': (C.c_int, [C.c_void_p, C.c_char_p, _f32p, _f32p]),
    'argsim_get_step': (C.c_int, [C.c_void_p, _i64p]),
    'argsim_set_step': (C.c_int, [C.c_void_p, C.c_int64]),
    'argsim_set_seed': (C.c_int, [C.c_void_p, C.c_uint64]),
    'argsim_train_step': (C.c_int, [C.c_void_p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, _u8p, _f32p, C.c_int64,
                                    C.c_int64, C.c_int64, C.POINTER(StepStats)]),
    'argsim_grad_step': (C.c_int, [C.c_void_p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, _u8p, _f32p, C.c_int64,
                                   C.c_int64, C.c_int64, C.POINTER(StepStats)]),
    'argsim_train_step_submit': (C.c_int, [C.c_void_p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, _u8p, _f32p, C.c_int64,
                                           C.c_int64, C.c_int64]),
    'argsim_train_step_wait': (C.c_int, [C.c_void_p, C.POINTER(StepStats)]),
    'argsim_set_global_rows': (C.c_int, [C.c_void_p, _i64p, C.c_int32]),
    'argsim_eval_step': (C.c_int, [C.c_void_p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, _f32p, _f32p, C.c_int64,
                                   _f32p, _i64p, _i32p]),
    'argsim_embed': (C.c_int, [C.c_void_p, _i32p, C.c_int32, C.c_int32, _f32p]),
    'argsim_decode_init': (C.c_int, [C.c_void_p, _f32p, C.c_int32, _f32p]),
    'argsim_decode_step': (C.c_int, [C.c_void_p, _i32p, C.c_int32, _f32p, _i32p]),
    'argsim_decode': (C.c_int, [C.c_void_p, _f32p, C.c_int32, C.c_int32, _i32p, _i32p]),
    'argsim_save': (C.c_int, [C.c_void_p, C.c_char_p]),
    'argsim_load': (C.c_int, [C.c_void_p, C.c_char_p]),
    'argsim_bench_resident': (C.c_int, [C.c_void_p, C.c_int32, _f32p]),
    'argsim_launch_count': (C.c_int, [C.c_void_p, _i64p]),
    'argsim_profiler': (C.c_int, [C.c_void_p, C.c_int32]),
    'argsim_last_timings': (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), _f32p]),
    'argsim_test_gemm': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f32p,
                                   _f32p, _f32p, C.c_float, C.c_int32, _f32p, _f32p]),
    'argsim_test_softmax_ce': (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, _f32p, _i32p, C.c_float, C.c_int32, _f32p,
                                         _f32p, _f32p, _i32p, C.POINTER(C.c_double)]),
    'argsim_test_ts_mma': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f32p, _f32p, _f32p, _i64p]),
    'argsim_bench_kernel': (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, _f32p, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double)]),
    'argsim_plan_batch': (C.c_int, [_i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _u8p, _i32p,
                                    _i32p, _i64p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_char_p, C.c_int32]),
    'argsim_schedule': (None, [C.c_int64, C.c_float, C.c_float, _f32p, _f32p, _f32p]),
}


def lib():
    """loads (building first if sources are newer) the shared library; raises if that fails."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if _build.needs_build():
            if os.path.exists(_build.NVCC):
                path = _build.build()
            elif not os.path.exists(path):
                raise RuntimeError('libargsim_b200.so is not built and nvcc is not available')
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _tokens(x, name):
    x = np.ascontiguousarray(x, dtype=np.int32)
    if x.ndim != 2:
        raise ValueError('%s must be an int32 matrix (batch, time), got shape %r' % (name, x.shape))
    return x


def schedule(step, accelerate=1e-4, learn_rate=1e-3):
    k, a, u = C.c_float(), C.c_float(), C.c_float()
    lib().argsim_schedule(int(step), accelerate, learn_rate, C.byref(k), C.byref(a), C.byref(u))
    return dict(rate_keepwd=np.float32(k.value), rate_anneal=np.float32(a.value), rate_update=np.float32(u.value))


def plan_batch(src, tgt, bos=2, eos=1, keep=None):
    """host-only index pipeline (no GPU needed); returns a dict of numpy int32 arrays."""
    src = _tokens(src, 'src')
    b, Ts = src.shape
    has_tgt = tgt is not None
    if has_tgt:
        tgt = _tokens(tgt, 'tgt')
        Tt = tgt.shape[1]
    else:
        Tt = 0
    if keep is not None:
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        assert keep.shape == (b, Tt)
    o = dict(len_src=np.zeros(b, np.int32), len_tgt=np.zeros(b, np.int32), ids_src=np.zeros(b * Ts, np.int32),
             lead=np.zeros(b * (Tt + 1), np.int32), gold=np.zeros(b * (Tt + 1), np.int32),
             ref_row=np.zeros(b * (Tt + 1), np.int32), enc_last=np.zeros(b, np.int32), perm_src=np.zeros(b, np.int32),
             perm_dec=np.zeros(b, np.int32))
    counts = np.zeros(4, np.int64)
    err = C.create_string_buffer(512)
    rc = lib().argsim_plan_batch(_p(src, _i32p), _p(tgt, _i32p) if has_tgt else None, b, Ts, Tt, bos, eos, _p(keep, _u8p),
                                 _p(o['len_src'], _i32p), _p(o['len_tgt'], _i32p), _p(counts, _i64p), _p(o['ids_src'], _i32p),
                                 _p(o['lead'], _i32p), _p(o['gold'], _i32p), _p(o['ref_row'], _i32p), _p(o['enc_last'], _i32p),
                                 _p(o['perm_src'], _i32p), _p(o['perm_dec'], _i32p), err, 512)
    if rc != 0:
        raise ValueError(err.value.decode())
    S, N = int(counts[0]), int(counts[1])
    o['ids_src'] = o['ids_src'][:S]
    for k in ('lead', 'gold', 'ref_row'):
        o[k] = o[k][:N]
    o.update(S=S, N=N, Tmax_src=int(counts[2]), Tmax_dec=int(counts[3]))
    return o


def nccl_unique_id():
    buf = (C.c_uint8 * 128)()
    if lib().argsim_nccl_unique_id(buf) != 0:
        raise RuntimeError(lib().argsim_last_error(None).decode())
    return bytes(buf)


def test_gemm(impl, A, B, a_mn=0, b_mn=0, bias=None, alpha=1.0, C0=None, device=0):
    """C = alpha * op(A) op(B)^T (+bias) (+C0); A given as (M,K) if a_mn==0 else (K,M); B likewise."""
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    M, K = (A.shape if not a_mn else A.shape[::-1])
    N, K2 = (B.shape if not b_mn else B.shape[::-1])
    assert K == K2
    out = np.zeros((M, N), np.float32) if C0 is None else np.ascontiguousarray(C0, np.float32).copy()
    bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
    ms = C.c_float()
    rc = lib().argsim_test_gemm(device, impl, M, N, K, a_mn, b_mn, _p(A, _f32p), _p(B, _f32p), _p(bias, _f32p), alpha,
                                0 if C0 is None else 1, _p(out, _f32p), C.byref(ms))
    if rc != 0:
        raise RuntimeError(lib().argsim_last_error(None).decode())
    return out, ms.value


def test_softmax_ce(logits, labels=None, gscale=1.0, bf16=True, write_grad=True, device=0):
    """runs the fused softmax-CE kernel once on host data; returns dict(grad, loss_samp, err_samp, pred, stats)."""
    logits = np.ascontiguousarray(logits, np.float32)
    n, V = logits.shape
    lab = None if labels is None else np.ascontiguousarray(labels, np.int32)
    grad = np.zeros_like(logits)
    loss = np.zeros(n, np.float32)
    err = np.zeros(n, np.float32)
    pred = np.zeros(n, np.int32)
    stats = (C.c_double * 2)()
    rc = lib().argsim_test_softmax_ce(device, int(bf16), n, V, _p(logits, _f32p), _p(lab, _i32p), gscale, int(write_grad),
                                      _p(grad, _f32p), _p(loss, _f32p), _p(err, _f32p), _p(pred, _i32p), stats)
    if rc != 0:
        raise RuntimeError(lib().argsim_last_error(None).decode())
    return dict(grad=grad, loss_samp=loss, err_samp=err, pred=pred, stats=np.array(stats[:]))


def test_ts_mma(A, B, nacc=1, device=0, want_cycles=False):
    """D = A . B^T through tensor memory (TS-form tcgen05.mma); A (128,K), B (N,K), operands rounded to bf16; the K/16
    instructions go round robin over `nacc` accumulators."""
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    assert A.shape[0] == 128 and A.shape[1] == B.shape[1]
    D = np.zeros((128, B.shape[0]), np.float32)
    cyc = np.zeros(1, np.int64)
    rc = lib().argsim_test_ts_mma(device, B.shape[0], A.shape[1], nacc, _p(A, _f32p), _p(B, _f32p), _p(D, _f32p), _p(cyc, _i64p))
    if rc != 0:
        raise RuntimeError(lib().argsim_last_error(None).decode())
    return (D, int(cyc[0])) if want_cycles else D


DEV_SIGNATURES = {
    'argsim_bench_exchange': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), _i32p]),
    'argsim_dev_last_error': (C.c_char_p, []),
}
_dev = None


def dev_lib():
    """libargsim_b200_dev.so (include/argsim_b200_dev.h): development microbenchmarks, not loaded by anything on the hot path."""
    global _dev
    if _dev is None:
        lib()                                   # builds both libraries when the sources are newer
        L = C.CDLL(_build.DEV_LIB)
        for name, (res, args) in DEV_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _dev = L
    return _dev


def bench_exchange(method, groups, rows, iters=2000, device=0):
    """cycles per 16-CTA all-gather round for one exchange mechanism (argsim_bench_exchange, development library)."""
    cyc = C.c_double()
    mc = np.zeros(1, np.int32)
    rc = dev_lib().argsim_bench_exchange(device, method, groups, rows, iters, C.byref(cyc), _p(mc, _i32p))
    if rc != 0:
        raise RuntimeError(dev_lib().argsim_dev_last_error().decode())
    return cyc.value, int(mc[0])


class Handle:
    """owns one argsim_handle (weights, Adam slots, step, RNG, device memory) -- the tf.Session analogue."""

    def __init__(self, dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, bidirectional=True, bidir_stacked=True,
                 attentive=False, logit_use_embed=True, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1,
                 precision=BF16, device=0, nranks=1, rank=0, nccl_id=None, flags=0, max_batch=0, max_len=0):
        self.L = lib()
        cfg = Config(dim_tgt, dim_emb, dim_rep, rnn_layers, int(bidirectional), int(bidir_stacked), int(attentive),
                     int(logit_use_embed), accelerate, learn_rate, bos, eos, precision, max_batch, max_len, device, nranks,
                     rank, (C.c_uint8 * 128)(*(nccl_id or bytes(128))), flags)
        self.cfg = cfg
        self.h = C.c_void_p()
        if self.L.argsim_create(C.byref(cfg), C.byref(self.h)) != 0:
            self.h = None
            raise RuntimeError('argsim_create failed: ' + self.L.argsim_last_error(None).decode())
        self.dim_rep, self.dim_emb, self.dim_tgt, self.rnn_layers = dim_rep, dim_emb, dim_tgt, rnn_layers
        self._shapes = None

    def close(self):
        if getattr(self, 'h', None):
            self.L.argsim_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.argsim_last_error(self.h).decode())

    # ---- parameters -------------------------------------------------------------------
    def param_shapes(self):
        if self._shapes is None:
            n = C.c_int32()
            self._ck(self.L.argsim_param_count(self.h, C.byref(n)))
            out = {}
            for i in range(n.value):
                name, rank, shape = C.c_char_p(), C.c_int32(), (C.c_int64 * 2)()
                self._ck(self.L.argsim_param_info(self.h, i, C.byref(name), C.byref(rank), shape))
                out[name.value.decode()] = tuple(shape[:rank.value])
            self._shapes = out
        return self._shapes

    def init_params(self, seed=0):
        self._ck(self.L.argsim_init_params(self.h, seed))

    def _get(self, fn, name):
        a = np.empty(self.param_shapes()[name], np.float32)
        self._ck(fn(self.h, name.encode(), _p(a, _f32p)))
        return a

    def get_param(self, name):
        return self._get(self.L.argsim_get_param, name)

    def get_grad(self, name):
        return self._get(self.L.argsim_get_grad, name)

    def set_param(self, name, value):
        a = np.ascontiguousarray(value, np.float32)
        if a.shape != self.param_shapes()[name]:
            raise ValueError('%s: expected shape %r, got %r' % (name, self.param_shapes()[name], a.shape))
        self._ck(self.L.argsim_set_param(self.h, name.encode(), _p(a, _f32p)))

    def set_params(self, params):
        for k, v in params.items():
            self.set_param(k, v)

    def get_params(self):
        return {k: self.get_param(k) for k in self.param_shapes()}

    def get_opt_state(self, name):
        m = np.empty(self.param_shapes()[name], np.float32)
        v = np.empty_like(m)
        self._ck(self.L.argsim_get_opt_state(self.h, name.encode(), _p(m, _f32p), _p(v, _f32p)))
        return m, v

    def set_opt_state(self, name, m, v):
        m = np.ascontiguousarray(m, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        self._ck(self.L.argsim_set_opt_state(self.h, name.encode(), _p(m, _f32p), _p(v, _f32p)))

    @property
    def step(self):
        s = C.c_int64()
        self._ck(self.L.argsim_get_step(self.h, C.byref(s)))
        return s.value

    @step.setter
    def step(self, v):
        self._ck(self.L.argsim_set_step(self.h, int(v)))

    def set_seed(self, seed):
        self._ck(self.L.argsim_set_seed(self.h, int(seed)))

    # ---- steps ------------------------------------------------------------------------
    def _step(self, fn, src, tgt, keep, eps, n_tokens_global, b_global, row0, stats=True, rows=None):
        src, tgt = _tokens(src, 'src'), _tokens(tgt, 'tgt')
        if src.shape[0] != tgt.shape[0]:
            raise ValueError('src and tgt must have the same number of rows')
        b = src.shape[0]
        if rows is not None:   # global row indices: RNG keying invariant to the sharding (argsim_set_global_rows)
            rows = np.ascontiguousarray(rows, np.int64)
            if rows.shape != (b,):
                raise ValueError('rows must hold one global row index per batch row')
            self._ck(self.L.argsim_set_global_rows(self.h, _p(rows, _i64p), b))
        if keep is not None:
            keep = np.ascontiguousarray(keep, np.uint8)
            if keep.shape != tgt.shape:
                raise ValueError('keep mask must have the shape of tgt')
        if eps is not None:
            eps = np.ascontiguousarray(eps, np.float32)
            if eps.shape != (b, self.dim_rep):
                raise ValueError('eps must be (batch, dim_rep)')
        if not stats:
            self._ck(fn(self.h, _p(src, _i32p), _p(tgt, _i32p), b, src.shape[1], tgt.shape[1], _p(keep, _u8p), _p(eps, _f32p),
                        n_tokens_global, b_global, row0))
            return None
        st = StepStats()
        self._ck(fn(self.h, _p(src, _i32p), _p(tgt, _i32p), b, src.shape[1], tgt.shape[1], _p(keep, _u8p), _p(eps, _f32p),
                    n_tokens_global, b_global, row0, C.byref(st)))
        return {k: getattr(st, k) for k, _ in StepStats._fields_}

    def train_step(self, src, tgt, keep=None, eps=None, n_tokens_global=0, b_global=0, row0=0, rows=None):
        return self._step(self.L.argsim_train_step, src, tgt, keep, eps, n_tokens_global, b_global, row0, rows=rows)

    def train_step_submit(self, src, tgt, keep=None, eps=None, n_tokens_global=0, b_global=0, row0=0, rows=None):
        """enqueue one training step and return (the arrays are not referenced afterwards); at most two un-waited steps."""
        self._step(self.L.argsim_train_step_submit, src, tgt, keep, eps, n_tokens_global, b_global, row0, stats=False, rows=rows)

    def train_step_wait(self):
        """statistics of the oldest un-waited step (blocks until it has finished on the device)."""
        st = StepStats()
        self._ck(self.L.argsim_train_step_wait(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in StepStats._fields_}

    def grad_step(self, src, tgt, keep=None, eps=None, n_tokens_global=0, b_global=0, row0=0, rows=None):
        return self._step(self.L.argsim_grad_step, src, tgt, keep, eps, n_tokens_global, b_global, row0, rows=rows)

    def eval_step(self, src, tgt, want_pred=False):
        src, tgt = _tokens(src, 'src'), _tokens(tgt, 'tgt')
        b = src.shape[0]
        cap = b * (tgt.shape[1] + 1)
        errt, lgen = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        lkld = np.zeros((b, self.dim_rep), np.float32)
        pred = np.zeros(cap, np.int32) if want_pred else None
        n = C.c_int64()
        self._ck(self.L.argsim_eval_step(self.h, _p(src, _i32p), _p(tgt, _i32p), b, src.shape[1], tgt.shape[1], _p(errt, _f32p),
                                         _p(lgen, _f32p), cap, _p(lkld, _f32p), C.byref(n), _p(pred, _i32p)))
        o = dict(errt_samp=errt[:n.value], loss_gen_samp=lgen[:n.value], loss_kld_samp=lkld)
        if want_pred:
            o['pred'] = pred[:n.value]
        return o

    def embed(self, src):
        src = _tokens(src, 'src')
        mu = np.empty((src.shape[0], self.dim_rep), np.float32)
        self._ck(self.L.argsim_embed(self.h, _p(src, _i32p), src.shape[0], src.shape[1], _p(mu, _f32p)))
        return mu

    def decode_init(self, z):
        z = np.ascontiguousarray(z, np.float32)
        b = z.shape[0]
        state = np.empty((self.rnn_layers, b, self.dim_emb), np.float32)
        self._ck(self.L.argsim_decode_init(self.h, _p(z, _f32p), b, _p(state, _f32p)))
        return state

    def decode_step(self, lead, state):
        lead = np.ascontiguousarray(lead, np.int32).reshape(-1)
        state = np.ascontiguousarray(state, np.float32).copy()
        pred = np.empty(lead.shape[0], np.int32)
        self._ck(self.L.argsim_decode_step(self.h, _p(lead, _i32p), lead.shape[0], _p(state, _f32p), _p(pred, _i32p)))
        return pred, state

    def decode(self, z, steps=256):
        """greedy decode of latent states, whole loop on the device: int32 (b, t), t <= steps (may be 0)."""
        z = np.ascontiguousarray(z, np.float32)
        b = z.shape[0]
        tok = np.zeros((b, steps), np.int32)
        t = np.zeros(1, np.int32)
        self._ck(self.L.argsim_decode(self.h, _p(z, _f32p), b, steps, _p(tok, _i32p), _p(t, _i32p)))
        return np.ascontiguousarray(tok[:, :int(t[0])])

    def save(self, path):
        self._ck(self.L.argsim_save(self.h, str(path).encode()))

    def load(self, path):
        self._ck(self.L.argsim_load(self.h, str(path).encode()))

    # ---- measurement ------------------------------------------------------------------
    def bench_resident(self, iters):
        ms = C.c_float()
        self._ck(self.L.argsim_bench_resident(self.h, iters, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.argsim_launch_count(self.h, C.byref(n)))
        return n.value

    def profiler(self, on):
        """cudaProfilerStart/Stop + NVTX ranges around the device programs (train.py --profile)."""
        self._ck(self.L.argsim_profiler(self.h, int(bool(on))))

    def last_timings(self):
        names = (C.c_char_p * 256)()
        ms = (C.c_float * 256)()
        n = self.L.argsim_last_timings(self.h, 256, names, ms)
        return {names[i].decode(): ms[i] for i in range(max(n, 0))}

    def bench_kernel(self, which, rows, iters=10):
        ms, by, fl = C.c_float(), C.c_double(), C.c_double()
        self._ck(self.L.argsim_bench_kernel(self.h, which.encode(), rows, iters, C.byref(ms), C.byref(by), C.byref(fl)))
        return ms.value, by.value, fl.value
