#!/bin/bash
mkdir -p gpurun_out
run() {
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err || tail -5 gpurun_out/q_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print('$1 ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['phases_ms'], d['last_step']['loss'])
PY
}
run "default  "
ARGSIM_DEC_EARLY=1 run "dec early"
run "default  "
ARGSIM_DEC_EARLY=1 run "dec early"
ARGSIM_DEC_EARLY=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
