#!/bin/bash
mkdir -p gpurun_out
for VAR in 26 58; do
export ARGSIM_GRU_VARIANT=$VAR
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m8_bench.json 2> gpurun_out/m8_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/m8_bench.json').read().strip().splitlines()[-1])
print('ARGSIM_GRU_VARIANT=$VAR ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'])
PY
done
export ARGSIM_GRU_VARIANT=58
ARGSIM_GRU_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/var.json 2> gpurun_out/var.err
grep gru_prof gpurun_out/var.err | grep "bwd_enc" | tail -2 | cut -c1-200
