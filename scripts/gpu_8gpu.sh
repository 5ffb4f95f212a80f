#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/gpu8_dp8.json 2> gpurun_out/gpu8_dp8.err
echo "dp8 rc=$?"; tail -2 gpurun_out/gpu8_dp8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload scaled --steps 4 --warmup 2 > gpurun_out/gpu8_scaled8.json 2> gpurun_out/gpu8_scaled8.err
echo "scaled8 rc=$?"; tail -2 gpurun_out/gpu8_scaled8.err
ATTENTIVE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/dp_check.py 2>&1 | grep -E "DP_CHECK|grad |stats|Error|error" | tail -16
python - <<'PY'
import json
d = json.loads(open('gpurun_out/gpu8_dp8.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp_check', d.get('dp_check'))
print(d['phases_ms'])
print('strong', d['strong_scaling']['ms_per_step'], d['strong_scaling']['value'])
print('embed', d['embed']['value'])
d = json.loads(open('gpurun_out/gpu8_scaled8.json').read().strip().splitlines()[-1])
print('scaled N', d['n_gpus'], 'ms', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), d['phases_ms'], d['step_tflops'])
PY
