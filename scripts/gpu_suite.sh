#!/bin/bash
mkdir -p gpurun_out
( time ARGSIM_TRAJ_STEPS=3 timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/suite_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/suite_tests.log
grep -E "^E |passed|failed|rc=|^real" gpurun_out/suite_tests.log | tail -6
python scripts/gpu_embed_prof.py 2>&1 | sed -n 2,3p
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/suite_bench.json 2>gpurun_out/suite_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/suite_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), '   embed', round(d['embed']['value']), 'seq/s', round(d['embed']['ms_per_batch'], 2), 'ms', d['embed']['roofline']['frac'], ';  strong b512', round(d['strong_scaling']['ms_per_step'], 2), 'ms')
PY
