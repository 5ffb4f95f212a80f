#!/bin/bash
mkdir -p gpurun_out
run() {
timeout 16 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c8_$2.json 2> gpurun_out/c8_$2.err || tail -2 gpurun_out/c8_$2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c8_$2.json').read().strip().splitlines()[-1])
    print('$1 ms', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'])
except Exception as e: print('$1 failed', e)
PY
}
ARGSIM_GRU_CHUNK=8 run "chunk8 " a
run "default" b
