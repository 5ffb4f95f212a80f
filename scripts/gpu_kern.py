import sys, os, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from argsim_b200 import _lib
cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
h = _lib.Handle(precision=_lib.BF16, **cfg)
h.init_params(0)
pk = json.load(open(os.path.join(R, 'MEASURED_PEAKS.json')))
out = {}
for which, rows in (('softmax_ce', 8872), ('softmax_ce', 65536), ('adam', 0), ('embed_gather', 17680), ('embed_gather', 262144), ('logits_gemm', 8872), ('logits_gemm', 65536)):
    ms, by, fl = h.bench_kernel(which, max(rows, 1), 10)
    r = dict(ms=ms, GBs=by / ms / 1e6, frac_hbm=by / ms / 1e6 / pk['hbm_gbs'])
    if fl:
        r.update(TFs=fl / ms / 1e9, frac_tensor_burst=fl / ms / 1e9 / pk['bf16_tflops'])
    out['%s@%d' % (which, rows)] = r
    print(which, rows, r, flush=True)
json.dump(out, open(os.path.join(R, 'gpurun_out', 'kernels_standalone.json'), 'w'), indent=1)
