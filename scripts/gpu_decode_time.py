"""greedy decode (src/model.py:204-219) at config.json dimensions: tokens/s of the device-resident loop, fp32 SIMT vs bf16
tensor-core step.  A bias on one token keeps the loop from stopping at all-eos so both run the full budget."""
import sys, os, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from argsim_b200 import _lib
cfg = dict(dim_tgt=8192, dim_emb=512, dim_rep=1024, rnn_layers=3, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1)
out = {}
for b in (16, 128, 512, 2048):
    z = np.random.default_rng(0).standard_normal((b, 1024)).astype(np.float32)
    for name, prec in (('fp32_simt', _lib.FP32_VALIDATE), ('bf16_tc', _lib.BF16)):
        h = _lib.Handle(precision=prec, **cfg)
        h.init_params(0)
        steps = 64
        h.decode(z, steps=steps)
        t0 = time.perf_counter(); tok = h.decode(z, steps=steps); dt = time.perf_counter() - t0
        out['%s_b%d' % (name, b)] = dict(ms_per_token_step=dt / max(tok.shape[1], 1) * 1e3, steps=int(tok.shape[1]),
                                         tokens_per_s=b * tok.shape[1] / dt)
        print(name, 'b', b, 'steps', tok.shape[1], 'ms/step %.3f' % (dt / max(tok.shape[1], 1) * 1e3), 'tokens/s %.0f' % (b * tok.shape[1] / dt), flush=True)
        h.close()
json.dump(out, open(os.path.join(R, 'gpurun_out', 'decode_time.json'), 'w'), indent=1)
