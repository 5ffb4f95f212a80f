#!/bin/bash
mkdir -p gpurun_out
for PW in 0 2 0 2; do
export ARGSIM_GRU_PAD_WAVE=$PW
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m19_bench.json 2> gpurun_out/m19_bench.err
python - <<PY
import json,os
try:
    d=json.loads(open('gpurun_out/m19_bench.json').read().strip().splitlines()[-1])
    print('pad_wave=$PW ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'])
except Exception as e:
    print('pad_wave=$PW FAILED', e); print(open('gpurun_out/m19_bench.err').read()[-600:])
PY
done
ARGSIM_GRU_PAD_WAVE=2 timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2
