"""torchrun --nproc-per-node N scripts/dp_check.py : N-GPU data parallel == 1-GPU on the same global batch
(losses, gradients after the bucketed NCCL all-reduce, parameters after Adam)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch
import torch.distributed as dist
from argsim_b200 import _lib, parallel
from argsim_b200.synth import synth_batch

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('cpu:gloo,cuda:nccl')
cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
if os.environ.get('ATTENTIVE'):   # src/model.py:136-145: its tensors share the latent affines' all-reduce bucket
    cfg['attentive'] = True
obj = [_lib.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
prec = _lib.BF16 if os.environ.get('PREC', 'bf16') == 'bf16' else _lib.FP32_VALIDATE
h = _lib.Handle(precision=prec, device=local, nranks=world, rank=rank, nccl_id=obj[0], **cfg)
h.init_params(0)          # same seed on every rank -> identical replicas
full = synth_batch(16 * world, 'iac', cfg['dim_tgt'], seed=1, cap=48)
rng = np.random.default_rng(2)
keep = (rng.random(full.shape) < 0.7).astype(np.uint8)
eps = rng.standard_normal((len(full), cfg['dim_rep'])).astype(np.float32)
s, t, rows, n_glob, b_glob = parallel.shard_batch(full, full, world, rank)
h.step = 7000
st = h.train_step(s, t, keep=keep[rows][:, :t.shape[1]], eps=eps[rows], n_tokens_global=n_glob, b_global=b_glob, row0=rank * len(rows))
names = ['embed/embedding', 'decode/out/kernel', 'decode/rnn/l1/R', 'latent/mu/kernel', 'encode/rnn2/bwd/W', 'encode/rnn1/fwd/bR']
if cfg.get('attentive'):
    names += ['encode/cata/k/kernel', 'encode/cata/q/bias', 'encode/cata/LayerNorm/gamma', 'encode/rnn3/bwd/R']
g = {k: h.get_grad(k) for k in names}
p = {k: h.get_param(k) for k in names}
# all ranks must hold identical reduced gradients and parameters
for k in names:
    a = torch.tensor(g[k]).cuda(); b = a.clone(); dist.broadcast(b, src=0)
    assert torch.equal(a, b), ('grad differs across ranks', k)
    a = torch.tensor(p[k]).cuda(); b = a.clone(); dist.broadcast(b, src=0)
    assert torch.equal(a, b), ('param differs across ranks', k)
if rank == 0:
    h1 = _lib.Handle(precision=prec, device=local, **cfg)
    h1.init_params(0)
    h1.step = 7000
    st1 = h1.train_step(full, full, keep=keep, eps=eps)
    print('DP stats ', {k: st[k] for k in ('loss', 'loss_gen', 'loss_kld', 'errt', 'n_tokens')})
    print('1GPU stats', {k: st1[k] for k in ('loss', 'loss_gen', 'loss_kld', 'errt', 'n_tokens')})
    tol = 2e-3 if prec == _lib.BF16 else 1e-5
    for k in ('loss', 'loss_gen', 'loss_kld', 'errt'):
        assert abs(st[k] - st1[k]) <= tol * max(abs(st1[k]), 1e-3), k
    for k in names:
        a, b = g[k].astype(np.float64).ravel(), h1.get_grad(k).astype(np.float64).ravel()
        cos = a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30)
        print('  grad %-22s cos %.6f ratio %.5f' % (k, cos, np.linalg.norm(a) / (np.linalg.norm(b) + 1e-30)))
        assert cos > (0.999 if prec == _lib.BF16 else 0.999999), k
    print('DP_CHECK_OK world=%d' % world)
dist.barrier()
dist.destroy_process_group()
