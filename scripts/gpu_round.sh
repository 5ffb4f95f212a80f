#!/bin/bash
# round-end style pass on one GPU: all GPU tests, smoke, both bench arms (default flags), embed workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke > gpurun_out/full_smoke.log 2>&1; tail -3 gpurun_out/full_smoke.log
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; tail -3 gpurun_out/full_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/full_bench.json').read().strip().splitlines()[-1])
print('OURS ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['clocks'])
print('roofline', d['roofline'])
print('cpu_baseline', d.get('cpu_baseline'))
PY
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/full_bench_ref.json 2> gpurun_out/full_bench_ref.err; tail -c 900 gpurun_out/full_bench_ref.json | cut -c1-900
