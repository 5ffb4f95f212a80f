#!/bin/bash
mkdir -p gpurun_out
run() {
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m20_bench.json 2> gpurun_out/m20_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/m20_bench.json').read().strip().splitlines()[-1])
print('$1 ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'])
PY
}
run "default            "
ARGSIM_ENC_SEG_FWD=1 run "fwd segmented 171  "
ARGSIM_ENC_SEG_FWD=1 ARGSIM_ENC_SEG=256 run "fwd+bwd seg 256    "
ARGSIM_ENC_SEG_FWD=1 ARGSIM_ENC_SEG=128 run "fwd+bwd seg 128    "
ARGSIM_ENC_SEG=128 run "bwd seg 128        "
ARGSIM_ENC_SEG=256 run "bwd seg 256        "
run "default            "
