#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q -k "persistent or config" 2>&1 | grep -v "^  \|^$\|^array\|^       " | tail -30 > gpurun_out/m3_tests.log; cat gpurun_out/m3_tests.log
ARGSIM_GRU_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err
grep gru_prof gpurun_out/prof_bench.err | tail -12 | cut -c1-160
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m3_bench.json 2> gpurun_out/m3_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/m3_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print({k:(v['ms_per_step']) for k,v in d['kernels'].items()})
print(d['phases_ms'])
PY
tail -3 gpurun_out/m3_bench.err
