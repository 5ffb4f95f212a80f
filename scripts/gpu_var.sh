#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x 2>&1 | tail -6
ARGSIM_DEC_SEG=0 ARGSIM_GRU_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/var.json 2> gpurun_out/var.err
grep "cluster_ok" gpurun_out/var.err | head -2
grep gru_prof gpurun_out/var.err | grep cycles | tail -12 | cut -c1-140 | awk 'NR%3==0'
tail -3 gpurun_out/var.err | cut -c1-300
for NC in 1 0; do
if [ $NC = 1 ]; then export ARGSIM_GRU_NO_CLUSTER=1; else unset ARGSIM_GRU_NO_CLUSTER; fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m3_bench.json 2> gpurun_out/m3_bench.err
python - <<'PY'
import json,os
d=json.loads(open('gpurun_out/m3_bench.json').read().strip().splitlines()[-1])
print('no_cluster', os.environ.get('ARGSIM_GRU_NO_CLUSTER'), 'ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')})
PY
done
