#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_golden.py -m gpu -q 2>&1 | grep -v "^  \|^$\|^array\|^       " | tail -30 > gpurun_out/m3_tests.log; cat gpurun_out/m3_tests.log
for SEG in 0 32 64 128; do
ARGSIM_DEC_SEG=$SEG timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m3_bench_$SEG.json 2> gpurun_out/m3_bench_$SEG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/m3_bench_$SEG.json').read().strip().splitlines()[-1])
print('SEG $SEG ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')})
PY
tail -2 gpurun_out/m3_bench_$SEG.err
done
