#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_train_driver.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/pipe_bench.json 2> gpurun_out/pipe_bench.err || tail -5 gpurun_out/pipe_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/pipe_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', d['e2e'])
PY
