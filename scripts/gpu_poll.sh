#!/bin/bash
# A/B of the forward recurrence's poll variants (ARGSIM_GRU_FWD2_OPT: 32 = two interleaved poll streams, 64 = first poll held back)
mkdir -p gpurun_out
run() {
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/poll_bench.json 2> gpurun_out/poll_bench.err || tail -3 gpurun_out/poll_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/poll_bench.json').read().strip().splitlines()[-1])
print('$1 ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')}, d['last_step']['loss'])
PY
}
run "default        "
for g in 400 600 800 1000 1200 1400; do
ARGSIM_GRU_FWD2_OPT=64 ARGSIM_GRU_POLL_GAP=$g run "held back $g"
done
run "default        "
