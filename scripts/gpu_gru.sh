#!/bin/bash
mkdir -p gpurun_out
for OPT in 0 12 16 0 v1; do
unset ARGSIM_GRU_FWD_V1; export ARGSIM_GRU_FWD2_OPT=$OPT
if [ $OPT = v1 ]; then export ARGSIM_GRU_FWD_V1=1; fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m9_bench.json 2> gpurun_out/m9_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/m9_bench.json').read().strip().splitlines()[-1])
print('opt=$OPT ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'])
PY
done
