#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m4_bench.json 2> gpurun_out/m4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/m4_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
print({k:(round(v['ms_per_step'],3), round(v['frac'],3)) for k,v in d['kernels'].items()})
print(d['phases_ms'])
PY
tail -3 gpurun_out/m4_bench.err
