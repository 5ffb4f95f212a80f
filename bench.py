#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: train sequences/sec of the
ELBO step: forward + backward + Adam) on synthetic IAC-shaped sentencepiece token batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|embed]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One rank per GPU.  Weak scaling: every rank owns 64 sequences of a seed-0 global batch of 64*N
(N=1 is BASELINE configs[1]; N=8 is configs[2], global batch 512), sharded balanced on length.
`value`  : device-timed (CUDA events inside the library, max over ranks) with the batch resident in HBM.
`e2e`    : the same K steps through the public C-ABI with HOST buffers (plan + H2D + step + D2H of every
           step's statistics inside the timed region), host-timed, issued the way the train driver does:
           argsim_train_step_submit(n+1) before argsim_train_step_wait(n); `e2e.blocking` is the same with
           one blocking argsim_train_step per step.
`--impl reference`: the reference's TF graph cannot run (no TensorFlow; CudnnGRU is GPU-only), so the
CPU arm is the torch-CPU port of the same graph (oracle/vae_torch.py) on all host threads: FULL-length steps of the
same 64-row batch (no truncation, no extrapolation), same `config` as this arm's line.
The default line also carries `embed` (BASELINE configs[3]: encoder-only mu of 4096 IBM-shaped rows through
argsim_embed with host buffers), `strong_scaling` (BASELINE configs[2] read as a fixed global batch of 512 over the N
ranks; at N = 1 that is the 1-GPU b = 512 number) and, for N > 1, `dp_check` (after the timed steps: the ranks' step
statistics and a checksum of every parameter tensor agree).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, bidirectional=True, bidir_stacked=True,
           attentive=False, logit_use_embed=True, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
PER_GPU = 64
METRIC = 'train sequences/sec (ELBO fwd+bwd+Adam)'


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sust=p['bf16_tflops_sustained'], src='measured')
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class Clocks:
    """samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith('active') for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def algo_flops(S, N, b, D=512, R=1024, V=8192, L=3):
    """forward GEMM flops (2MNK), padding excluded; training = 3x forward."""
    H = D
    enc = (2 * 2 * 3 * H * (D + H) + 2 * (L - 1) * 2 * 3 * H * (2 * H + H)) * S
    dec = (L * 2 * 3 * H * (D + H) + 2 * D * D) * N
    voc = 2 * D * V * N
    lat = (2 * 2 * (2 * H) * R + 2 * R * D) * b
    return dict(enc=enc, dec=dec, voc=voc, lat=lat, fwd=enc + dec + voc + lat, train=3 * (enc + dec + voc + lat),
                rec_fwd=2 * 3 * H * H * (2 * L * S + L * N))


def dist_setup(n):
    """returns (rank, world, local_rank, torch.distributed or None)"""
    if n <= 1 and 'RANK' not in os.environ:
        return 0, 1, 0, None
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', rank))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend='cpu:gloo,cuda:nccl' if torch.cuda.is_available() else 'gloo')
        return rank, world, local, dist
    return 0, 1, 0, None


def workload_config(world, full=None, S_glob=None, N_glob=None):
    """`config` of the JSON line: the SAME dict for both arms (the reference arm runs rank 0's N=1 workload)."""
    gb = PER_GPU * world
    if full is None:
        from argsim_b200.synth import synth_batch
        full = synth_batch(gb, 'iac', CFG['dim_tgt'], seed=0)
        S_glob = int((full != 1).sum())
        N_glob = int(((full != 1).sum(1) + 1).sum())
    return dict(workload='config.json VAE (V=8192 D=512 R=1024 L=3), IAC-shaped synthetic sentencepiece batch, '
                         'seed 0, %d sequences per GPU (BASELINE configs[%d])' % (PER_GPU, 1 if world == 1 else 2),
                global_batch=gb, src_tokens=S_glob, tgt_rows=N_glob, max_len=int(full.shape[1]),
                parallelism='dp%d' % world, l2='no flush needed: one step streams >1 GB of activations and 0.68 GB of '
                'Adam state, far above the 126 MB L2')


class CpuPort:
    """torch-CPU port of the reference graph (oracle/vae_torch.py) on the C1 batch.  The reference pads every
    row to the batch maximum and runs cuDNN over all padded steps, so the cost of a step is proportional to the
    number of TIME STEPS: a bounded sample is the same 64 rows truncated to their first T' tokens, and the
    full-length rate is extrapolated as 64 / (t_sample * 512 / T')."""

    def __init__(self):
        import torch
        from argsim_b200.synth import synth_batch
        from oracle import vae_oracle as O
        from oracle import vae_torch as T
        self.T, self.torch = T, torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.full = synth_batch(PER_GPU, 'iac', CFG['dim_tgt'], seed=0)
        self.tmax = self.full.shape[1]
        self.P = T.to_torch(O.init_params(CFG, seed=0, dtype=np.float32), torch.float32, requires_grad=True)
        self.M = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.V = {k: torch.zeros_like(v) for k, v in self.P.items()}
        rng = np.random.default_rng(0)
        self.keep = (rng.random((self.tmax, PER_GPU)) < 0.5).astype(np.int64)
        self.eps = rng.standard_normal((PER_GPU, CFG['dim_rep'])).astype(np.float32)
        self.it = 0

    def step(self, tprime):
        sub = self.full[:, :tprime].copy()
        t0 = time.perf_counter()
        self.T.train_step(self.P, self.M, self.V, CFG, sub, sub, self.it, self.keep[:tprime], self.eps)
        self.it += 1
        return time.perf_counter() - t0

    def pick(self, nsteps, budget_s):
        """full length unless a budget is given (ARGSIM_REF_BUDGET_S): then the largest power-of-two truncation whose
        nsteps steps fit it (calibrated on T'=8); a truncated run is labelled as such and never extrapolated silently"""
        if not budget_s:
            return self.tmax
        self.step(8)
        t8 = self.step(8)
        tp = 8
        while tp * 2 <= self.tmax and nsteps * t8 * (tp * 2 / 8.0) <= budget_s:
            tp *= 2
        return self.tmax if tp * 2 > self.tmax else tp


def reference_arm(args, rank, world):
    """--impl reference: the CPU arm on the box's host cores, all threads, bounded sample per step."""
    if rank != 0:
        return
    port = CpuPort()
    tp = port.pick(args.steps + args.warmup, float(os.environ.get('ARGSIM_REF_BUDGET_S', 0)))
    times = [port.step(tp) for _ in range(args.warmup + args.steps)][args.warmup:]
    ms = 1e3 * float(np.mean(times))
    full = tp == port.tmax
    ms_full = ms * port.tmax / tp
    val = PER_GPU / (ms_full / 1e3)
    if full:
        sample = ('one step = the whole 64-row C1 batch, all %d padded time steps, forward + backward + TF-form Adam in torch-CPU '
                  'fp32 (library GRU + MKL GEMMs); %d timed steps after %d warm-up steps, nothing extrapolated'
                  % (port.tmax, args.steps, args.warmup))
    else:
        sample = ('EXTRAPOLATED (ARGSIM_REF_BUDGET_S set): the 64-row C1 batch truncated to its first %d of %d time steps per '
                  'step, value scaled x%d/%d' % (tp, port.tmax, tp, port.tmax))
    line = dict(metric=METRIC, value=val, unit='sequences/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_full, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference', config=workload_config(world),
                note='the TF reference cannot run (no TensorFlow; CudnnGRU has no CPU kernel): torch-CPU port of the same padded '
                     'graph (oracle/vae_torch.py); rank 0 runs the per-GPU workload of the other arm (64 sequences), '
                     'sample_time_steps=%d of %d' % (tp, port.tmax),
                cpu_baseline=dict(value=val, unit='sequences/s', cores=port.cores, kind='port', sample=sample),
                e2e=dict(value=val, unit='sequences/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """rank 0, N=1: the same port timed for ~20 s next to the GPU number."""
    port = CpuPort()
    times = [port.step(port.tmax) for _ in range(3)][1:]
    ms_full = 1e3 * float(np.mean(times))
    return dict(value=PER_GPU / (ms_full / 1e3), unit='sequences/s', cores=port.cores, kind='port',
                sample='2 timed FULL steps (after 1 warm-up) of the same 64-row C1 batch, all %d padded time steps, forward + '
                       'backward + TF-form Adam in torch-CPU fp32 (TF reference not runnable); nothing extrapolated' % port.tmax)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    ap.add_argument('--workload', default='train', choices=('train', 'embed', 'scaled'))
    ap.add_argument('--precision', default='bf16', choices=('bf16', 'fp32'))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the embed / strong_scaling sub-records')
    ap.add_argument('--generic-gru', action='store_true', help='bf16 GEMMs with the per-step generic GRU (debug)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank, world, local, dist = dist_setup(args.gpus)
    if args.impl == 'reference':
        reference_arm(args, rank, world)
        if dist:
            dist.destroy_process_group()
        return

    import torch  # device plumbing only: barrier / max over ranks
    from argsim_b200 import _lib, parallel
    from argsim_b200.synth import synth_batch
    if not torch.cuda.is_available():
        sys.exit('bench.py needs a CUDA device (no CPU fallback)')
    pk = peaks()
    nccl_id = None
    if world > 1:
        obj = [_lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]
    flags = _lib.FLAG_KERNEL_TIMERS | (_lib.FLAG_GENERIC_GRU if args.generic_gru else 0)
    prec = _lib.BF16 if args.precision == 'bf16' else _lib.FP32_VALIDATE
    cfg4 = dict(CFG, dim_tgt=32768, dim_emb=2048, dim_rep=4096)     # BASELINE configs[4]
    h = _lib.Handle(precision=prec, device=local, nranks=world, rank=rank, nccl_id=nccl_id, flags=flags,
                    **(cfg4 if args.workload == 'scaled' else CFG))
    h.init_params(0)
    h.set_seed(0)

    def barrier():
        if dist:
            t = torch.zeros(1, device='cuda')
            dist.all_reduce(t)
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == 'scaled':
        # BASELINE configs[4]: scaled VAE (4x hidden width, 32k vocabulary, every row 512 tokens), data parallel, 64 rows per GPU
        full4 = synth_batch(PER_GPU * world, 'full', cfg4['dim_tgt'], seed=0, cap=512)
        s4, t4, rows4, ntok4, b4 = parallel.shard_batch(full4, full4, world, rank)
        kw4 = dict(n_tokens_global=ntok4, b_global=b4, rows=rows4) if world > 1 else {}
        for _ in range(max(2, min(args.warmup, 3))):
            st4 = h.train_step(s4, t4, **kw4)
        barrier()
        clk = Clocks(local)
        k4 = max(2, min(args.steps, 5))
        ms4 = max_over_ranks(h.bench_resident(k4))
        clocks4 = clk.stop()
        tm4 = h.last_timings()
        barrier()
        if rank == 0:
            S4 = int((full4 != 1).sum())
            H4, V4, R4 = 2048, 32768, 4096
            fl4 = 3.0 * ((2 * 2 * 3 * H4 * (H4 + H4) + 2 * 2 * 2 * 3 * H4 * (2 * H4 + H4)) * S4 + (3 * 2 * 3 * H4 * (H4 + H4) + 2 * H4 * H4) * ntok4
                         + 2 * H4 * V4 * ntok4 + (2 * 2 * (2 * H4) * R4 + 2 * R4 * H4) * b4)
            print(json.dumps(dict(metric=METRIC + ', scaled config', value=PER_GPU * world / (ms4 / 1e3), unit='sequences/s', n_gpus=world, steps=k4,
                                  warmup=3, ms_per_step=ms4, higher_is_better=True, scaling='weak', vs_baseline=None, dtype=args.precision,
                                  data='synthetic',
                                  config=dict(workload='scaled VAE (V=32768 D=2048 R=4096 L=3), every row 512 tokens, 64 sequences per GPU '
                                                       '(BASELINE configs[4])', global_batch=PER_GPU * world, src_tokens=S4, tgt_rows=int(ntok4),
                                              parallelism='dp%d' % world),
                                  clocks=clocks4, phases_ms={k: round(v, 3) for k, v in tm4.items() if not k.startswith('k:')},
                                  kernels_ms={k[2:]: round(v, 3) for k, v in tm4.items() if k.startswith('k:') and '#' not in k},
                                  step_tflops=round(fl4 / (ms4 * 1e-3) / 1e12, 2), step_frac_of_tensor_peak=round(fl4 / (ms4 * 1e-3) / 1e12 / pk['tf_sust'] / world, 4),
                                  last_step=dict(loss=st4['loss'], loss_gen=st4['loss_gen'], loss_kld=st4['loss_kld']))), flush=True)
        if dist:
            dist.destroy_process_group()
        return

    if args.workload == 'embed':
        data = synth_batch(4096 * world, 'ibm', CFG['dim_tgt'], seed=0)
        mine = data[rank::world]
        mine = np.ascontiguousarray(mine[:, :int((mine != 1).sum(1).max())])
        for _ in range(args.warmup):
            h.embed(mine)
        barrier()
        clk = Clocks(local)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            h.embed(mine)
        dt = time.perf_counter() - t0
        barrier()
        ms = max_over_ranks(1e3 * dt / args.steps)
        if rank == 0:
            S = int((data != 1).sum())
            print(json.dumps(dict(metric='embed sequences/sec (encoder mu)', value=4096 * world / (ms / 1e3), unit='sequences/s',
                                  n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                                  scaling='weak', vs_baseline=None, dtype=args.precision, data='synthetic',
                                  config=dict(workload='encoder-only mu embedding, batch 4096 per GPU, IBM-shaped lengths (seed 0)',
                                              tokens=S), clocks=clk.stop(),
                                  e2e=dict(value=4096 * world / (ms / 1e3), unit='sequences/s',
                                           h2d_bytes_per_step=int(mine.nbytes), d2h_bytes_per_step=4096 * 1024 * 4))), flush=True)
        return

    gb = PER_GPU * world
    full = synth_batch(gb, 'iac', CFG['dim_tgt'], seed=0)
    src, tgt, rows, n_tok_glob, b_glob = parallel.shard_batch(full, full, world, rank)
    # rows = this rank's row indices in the global batch: they key the Philox streams (word dropout, eps), so the
    # un-injected randomness of the step does not depend on the number of ranks
    kw = dict(n_tokens_global=n_tok_glob, b_global=b_glob, rows=rows) if world > 1 else {}
    S_glob = int((full != 1).sum())
    N_glob = n_tok_glob

    # ---- e2e first (also serves as warm-up of the resident run): host buffers -> C ABI -> stats back.
    # (1) blocking calls, one argsim_train_step per step; (2) the form the train driver uses (Session.run(train_step),
    # src/train.py:118 fetches nothing but the op): argsim_train_step_submit(n+1) before argsim_train_step_wait(n), so
    # the host plan + H2D of a step overlap the device's previous step. Every step's inputs go host -> device and every
    # step's statistics come back inside the timed region in both.
    for _ in range(args.warmup):
        st = h.train_step(src, tgt, **kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = h.train_step(src, tgt, **kw)
    e2e_blk_dt = time.perf_counter() - t0
    barrier()
    e2e_blk_ms = max_over_ranks(1e3 * e2e_blk_dt / args.steps)
    t0 = time.perf_counter()
    h.train_step_submit(src, tgt, **kw)
    for _ in range(args.steps - 1):
        h.train_step_submit(src, tgt, **kw)
        st = h.train_step_wait()
    st = h.train_step_wait()
    e2e_dt = time.perf_counter() - t0
    barrier()
    e2e_ms = max_over_ranks(1e3 * e2e_dt / args.steps)
    plan = _lib.plan_batch(src, tgt)
    h2d = 4 * (plan['S'] + 2 * plan['N'] + 2 * len(src) + 2 * (plan['Tmax_src'] + plan['Tmax_dec']) + 2)
    d2h = 32

    # ---- device-resident timed region: exactly K steps, CUDA events on the library's stream
    for _ in range(2):
        h.bench_resident(1)
    barrier()
    clk = Clocks(local)
    l0 = h.launch_count()
    ms_local = h.bench_resident(args.steps)
    launches = h.launch_count() - l0
    barrier()
    clocks = clk.stop()
    ms = max_over_ranks(ms_local)
    tm = h.last_timings()   # phases + k: timers of the LAST step of the timed region

    # ---- dp_check (N > 1): the data-parallel invariant after the timed steps, with the schedule that was timed --
    # every rank's step statistics are the global ones and every parameter tensor is bit-identical across ranks
    dp_check = None
    if dist:
        import zlib
        sums = []
        for name in sorted(h.param_shapes()):
            a = h.get_param(name)
            sums.append(float(zlib.crc32(a.tobytes())))
            sums.append(float(np.abs(a.astype(np.float64)).sum()))
        mine_v = torch.tensor([st['loss'], st['loss_gen'], st['loss_kld'], float(st['step'])] + sums, device='cuda', dtype=torch.float64)
        lo, hi = mine_v.clone(), mine_v.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        bad = int((lo != hi).sum().item())
        finite = bool(torch.isfinite(mine_v).all().item())
        dp_check = 'ok' if bad == 0 and finite else 'FAILED: %d of %d values differ across ranks%s' % (bad, mine_v.numel(), '' if finite else ', non-finite')

    # ---- strong scaling: BASELINE configs[2] as a FIXED global batch of 512 dealt over the N ranks (N = 1: b = 512 on one GPU)
    strong = None
    if args.workload == 'train' and not args.no_extra:
        full512 = synth_batch(512, 'iac', CFG['dim_tgt'], seed=0)
        s5, t5, rows5, ntok5, b5 = parallel.shard_batch(full512, full512, world, rank)
        kw5 = dict(n_tokens_global=ntok5, b_global=b5, rows=rows5) if world > 1 else {}
        for _ in range(3):
            st5 = h.train_step(s5, t5, **kw5)
        barrier()
        ms5 = max_over_ranks(h.bench_resident(max(3, min(args.steps, 10))))
        tm5 = h.last_timings()
        strong = dict(global_batch=512, rows_per_gpu=int(len(s5)), n_gpus=world, ms_per_step=ms5, value=512 / (ms5 / 1e3), unit='sequences/s',
                      src_tokens=int((full512 != 1).sum()), tgt_rows=int(ntok5),
                      phases_ms={k: round(v, 4) for k, v in tm5.items() if not k.startswith('k:')},
                      note='same model and data generator, global batch fixed at 512 sequences; device-timed, resident batch')

    # ---- embed: BASELINE configs[3] / metric "embed seq/sec": mu of 4096 IBM-shaped rows per GPU through argsim_embed
    embed = None
    if args.workload == 'train' and not args.no_extra:
        data = synth_batch(4096, 'ibm', CFG['dim_tgt'], seed=rank)
        data = np.ascontiguousarray(data[:, :int((data != 1).sum(1).max())])
        for _ in range(2):
            h.embed(data)
        barrier()
        clk_e = Clocks(local)
        t0 = time.perf_counter()
        ne = 5
        for _ in range(ne):
            mu = h.embed(data)
        dt = time.perf_counter() - t0
        barrier()
        ms_e = max_over_ranks(1e3 * dt / ne)
        S_e = int((data != 1).sum())
        H_ = CFG['dim_emb']
        fl_e = (2 * 2 * 3 * H_ * (H_ + H_) + 2 * 2 * 2 * 3 * H_ * (2 * H_ + H_)) * S_e + 2 * (2 * H_) * CFG['dim_rep'] * 4096
        embed = dict(metric='embed sequences/sec (encoder mu)', value=4096 * world / (ms_e / 1e3), unit='sequences/s', ms_per_batch=ms_e,
                     batch_per_gpu=4096, tokens_per_gpu=S_e, h2d_bytes_per_step=int(data.nbytes), d2h_bytes_per_step=int(mu.nbytes),
                     timer='host perf_counter around %d argsim_embed calls with host buffers (plan + H2D + encoder + D2H of mu)' % ne,
                     roofline=dict(bound='tensor', achieved=round(fl_e / (ms_e * 1e-3) / 1e12, 3), peak=pk['tf_sust'], unit='TFLOP/s',
                                   frac=round(fl_e / (ms_e * 1e-3) / 1e12 / pk['tf_sust'], 5),
                                   note='algorithmic encoder flops (SURVEY 8d: 25.166 MFLOP per source token) over the e2e time'),
                     clocks=clk_e.stop(), finite=bool(np.isfinite(mu).all()))

    # ---- decode: greedy generation (src/model.py:204-219) of 128 latent rows, 64 token steps, through argsim_decode
    decode = None
    if args.workload == 'train' and not args.no_extra:   # every rank runs it (ranks stay in step); rank 0's number is reported
        zz = np.random.default_rng(7).standard_normal((128, CFG['dim_rep'])).astype(np.float32)
        h.decode(zz, steps=64)
        t0 = time.perf_counter()
        tok = h.decode(zz, steps=64)
        dt = time.perf_counter() - t0
        decode = dict(metric='greedy decode tokens/sec', value=tok.size / dt if tok.size else 0.0, unit='tokens/s', rows=128,
                      steps=int(tok.shape[1]), ms_per_token_step=1e3 * dt / max(int(tok.shape[1]), 1),
                      timer='host perf_counter around one argsim_decode call (z H2D, device-resident loop, tokens D2H)',
                      note='one GPU (rank 0 reported); weights of the timed run, so the loop may stop early at all-eos')

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    value = gb / (ms / 1e3)
    fl = algo_flops(S_glob, N_glob, gb)
    # dominant kernel = largest k: timer of the last timed step
    kt = {k[2:]: v for k, v in tm.items() if k.startswith('k:') and '#' not in k}
    kn = {k[2:-2]: v for k, v in tm.items() if k.startswith('k:') and k.endswith('#n')}
    kg = {k[2:-6]: v for k, v in tm.items() if k.startswith('k:') and k.endswith('#gflop')}   # GEMM groups: flops counted by the library
    phases = {k: round(v, 4) for k, v in tm.items() if not k.startswith('k:')}
    plan_l = _lib.plan_batch(src, tgt)
    S_l, N_l = plan_l['S'], plan_l['N']
    H = CFG['dim_emb']
    kalgo = {  # algorithmic work of one rank's step per kernel group (flops or bytes)
        'gru_fwd_enc': ('tensor', 2 * 3 * H * H * 2 * 3 * S_l), 'gru_bwd_enc': ('tensor', 2 * 3 * H * H * 2 * 3 * S_l),
        'gru_fwd_dec': ('tensor', 2 * 3 * H * H * 3 * N_l), 'gru_bwd_dec': ('tensor', 2 * 3 * H * H * 3 * N_l),
        'logits_gemm': ('tensor', 2 * H * CFG['dim_tgt'] * N_l), 'logits_wgrad': ('tensor', 2 * H * CFG['dim_tgt'] * N_l),
        'logits_dgrad': ('tensor', 2 * H * CFG['dim_tgt'] * N_l),
        'softmax_ce': ('hbm', (4 if args.precision == 'bf16' else 8) * N_l * CFG['dim_tgt']),
        'adam': ('hbm', 24410112 * (30 if args.precision == 'bf16' else 28)),
    }
    for k, gfl in kg.items():
        kalgo[k] = ('tensor', gfl * 1e9)
    if 'logits_wgrad_side' in kt:
        kalgo['logits_wgrad_side'] = kalgo['logits_wgrad']
    if 'adam_side' in kt:   # single GPU: all parameters in front of encode/rnn1 are updated on the side stream under the last recurrence
        n_tail = 2 * (3 * H * CFG['dim_emb'] + 3 * H * H + 6 * H) + CFG['dim_tgt'] * CFG['dim_emb']
        per = 30 if args.precision == 'bf16' else 28
        kalgo['adam'] = ('hbm', n_tail * per)
        kalgo['adam_side'] = ('hbm', (24410112 - n_tail) * per)
    kernels = {}
    for k, t in kt.items():
        if k in kalgo and t > 0:
            bound, work = kalgo[k]
            ach = work / (t * 1e-3) / (1e9 if bound == 'hbm' else 1e12)
            peak = pk['hbm'] if bound == 'hbm' else pk['tf_sust']
            kernels[k] = dict(bound=bound, ms_per_step=round(t, 4), launches_per_step=int(kn.get(k, 0)), achieved=round(ach, 3),
                              peak=peak, frac=round(ach / peak, 5), share_of_step=round(t / ms, 4))
            if k.endswith('_side'):   # low-priority side stream: runs on the SMs the concurrent recurrence launches leave free
                kernels[k]['overlapped'] = True
                kernels[k]['note'] = ('off the serial chain, sharing the chip with the recurrence kernels: the duration is stretched by '
                                      'design and frac is not a kernel-quality figure (stand-alone: profiles/r1_kernels_standalone.json)')
    dom = max((k for k in kernels if not k.endswith('_side')), key=lambda k: kernels[k]['ms_per_step']) if kernels else None
    roofline = None
    if dom:
        d = kernels[dom]
        traffic = None
        try:   # DRAM bytes per launch from the committed ncu --set full captures of this round (profiles/r2_traffic.json)
            with open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')) as f:
                tr = json.load(f)
            tkey = [k for k in tr if ('k_gru_mma_bwd' in k if dom.startswith('gru_bwd') else 'k_gru_mma_fwd2' in k if dom.startswith('gru_fwd') else False)]
            if tkey:
                traffic = tr[tkey[0]]['dram_bytes_per_launch']
        except Exception:
            pass
        roofline = dict(kernel=dom, bound=d['bound'], achieved=d['achieved'], peak=d['peak'],
                        unit='GB/s' if d['bound'] == 'hbm' else 'TFLOP/s', frac=d['frac'], traffic=traffic,
                        peak_source='%s (%s)' % (pk['src'], 'hbm_gbs' if d['bound'] == 'hbm' else 'bf16_tflops_sustained: timed inside a long step'),
                        note='avg launch duration from CUDA events on the launching stream inside the timed region; '
                             'the recurrence is bound by serial-step latency (512 dependent steps per launch, one 16-CTA '
                             'exchange through L2 each: >= 1270 cycles, DESIGN.md section 5), not by the tensor pipe; '
                             'traffic = dram bytes per launch of the same kernel on decoder segment launches (profiles/r2_gru_*_ncu.md, r2_traffic.json)')
        if dom.startswith('gru_'):
            # what actually bounds these kernels: serial steps x per-step latency.  Encoder: one launch (or segment chain)
            # per layer, Tmax_src dependent steps each; decoder: a 3-layer wavefront over Tmax_dec steps.
            enc = dom.endswith('_enc')
            serial = 3 * plan_l['Tmax_src'] if enc else plan_l['Tmax_dec'] + 2 * 64
            mhz = float(clocks.get('sm_mhz') or 1965.0)
            cyc = d['ms_per_step'] * 1e-3 / serial * mhz * 1e6
            roofline['latency'] = dict(serial_steps_per_training_step=int(serial), cycles_per_serial_step=round(cyc, 1),
                                       exchange_floor_cycles=950, frac_of_exchange_floor=round(950.0 / cyc, 4),
                                       note='one 16-CTA all-gather / reduce-scatter through L2 per serial step: 950 cycles measured '
                                            'for the bare exchange in the placement the launches use (argsim_bench_exchange, '
                                            'profiles/r1_exchange_xbench.json); the rest of a step is HMMA issue, one block barrier '
                                            'and the gate math (DESIGN.md section 5)')
    line = dict(metric=METRIC, value=value, unit='sequences/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='bf16' if args.precision == 'bf16' else 'f32', data='synthetic',
                config=workload_config(world, full, S_glob, N_glob),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=gb / (e2e_ms / 1e3), unit='sequences/s', ms_per_step=e2e_ms, h2d_bytes_per_step=int(h2d),
                         d2h_bytes_per_step=d2h,
                         timer='host perf_counter around K steps issued as argsim_train_step_submit(n+1) / argsim_train_step_wait(n) '
                               '(the train driver\'s form); statistics of every step read back',
                         blocking=dict(value=gb / (e2e_blk_ms / 1e3), ms_per_step=e2e_blk_ms,
                                       timer='host perf_counter around K blocking argsim_train_step calls')),
                roofline=roofline, kernels=kernels, phases_ms=phases,
                step_tflops=round(fl['train'] / (ms * 1e-3) / 1e12, 3), step_frac_of_tensor_peak=round(fl['train'] / (ms * 1e-3) / 1e12 / pk['tf_sust'], 5),
                last_step=dict(loss=st['loss'], loss_gen=st['loss_gen'], loss_kld=st['loss_kld']))
    if embed:
        line['embed'] = embed
    if decode:
        line['decode'] = decode
    if strong:
        line['strong_scaling'] = strong
    if dp_check:
        line['dp_check'] = dp_check
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline_leg()
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
