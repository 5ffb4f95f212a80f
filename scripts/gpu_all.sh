#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/m7_bench.json 2> gpurun_out/m7_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/m7_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
for k,v in sorted(d['kernels'].items()): print('  %-22s %7.3f ms  n=%d  %8.1f %s  frac %.3f' % (k, v['ms_per_step'], v['launches_per_step'], v['achieved'], 'GB/s' if v['bound']=='hbm' else 'TF/s', v['frac']))
print(d['phases_ms'])
PY
tail -3 gpurun_out/m7_bench.err
