#!/bin/bash
mkdir -p gpurun_out
for SEG in 64 48 32 96 40; do
export ARGSIM_DEC_SEG=$SEG
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m15_bench.json 2> gpurun_out/m15_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/m15_bench.json').read().strip().splitlines()[-1])
print('seg=$SEG ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')})
PY
done
