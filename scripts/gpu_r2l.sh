#!/bin/bash
mkdir -p gpurun_out
( time ARGSIM_TRAJ_STEPS=3 timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 ) > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2l_tests.log
grep -E "^E |passed|failed|rc=|^real" gpurun_out/r2l_tests.log | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'blocking', round(d['e2e']['blocking']['ms_per_step'], 3))
for k, v in sorted(d['kernels'].items()):
    print('  %-28s %8.4f ms  frac %.3f' % (k, v['ms_per_step'], v['frac']))
print('   embed', round(d['embed']['value']), 'seq/s;  strong b512', round(d['strong_scaling']['ms_per_step'], 2), 'ms')
PY
