#!/bin/bash
mkdir -p gpurun_out
for R in 16 24 32 40 16; do
export ARGSIM_SMALL8_ROWS=$R
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m14_bench.json 2> gpurun_out/m14_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/m14_bench.json').read().strip().splitlines()[-1])
print('small8_rows=$R ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'])
PY
done
