#!/bin/bash
# A/B: weight-gradient GEMMs / early Adam on the low-priority side stream (default) vs in line on the main stream
mkdir -p gpurun_out
run() {
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/wg_bench.json 2> gpurun_out/wg_bench.err || tail -3 gpurun_out/wg_bench.err
python - <<PY
import json,os
d=json.loads(open('gpurun_out/wg_bench.json').read().strip().splitlines()[-1])
print('$1 ms_per_step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru') or 'side' in k or 'adam' in k}, d['phases_ms'], d['last_step']['loss'])
PY
}
run "default (per-segment wgrad of rnn1)"
ARGSIM_NO_SEG_WGRAD=1 run "whole-layer wgrad of rnn1          "
run "default (per-segment wgrad of rnn1)"
ARGSIM_NO_SEG_WGRAD=1 run "whole-layer wgrad of rnn1          "
ARGSIM_WGRAD_OVERLAP=0 run "all in line                        "
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
