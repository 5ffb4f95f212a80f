#!/bin/bash
mkdir -p gpurun_out
ARGSIM_DEC_SEG=0 ARGSIM_GRU_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/var.json 2> gpurun_out/var.err
grep gru_prof gpurun_out/var.err | tail -12 | cut -c1-140 | awk 'NR%3==0'
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m3_bench.json 2> gpurun_out/m3_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/m3_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')})
print(d['phases_ms'])
PY
