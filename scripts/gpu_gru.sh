#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2 3; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m17_bench.json 2> gpurun_out/m17_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/m17_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'], d['gpu_launches'])
PY
done
