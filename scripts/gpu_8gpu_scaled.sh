#!/bin/bash
# BASELINE configs[4] on 8 GPUs only (the default workload's 8-GPU line comes from scripts/gpu_8gpu.sh)
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload scaled --steps 4 --warmup 2 > gpurun_out/gpu8_scaled8.json 2> gpurun_out/gpu8_scaled8.err
echo "scaled8 rc=$?"; tail -2 gpurun_out/gpu8_scaled8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/gpu8_scaled8.json').read().strip().splitlines()[-1])
print('scaled N', d['n_gpus'], 'ms', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), d['phases_ms'], d['step_tflops'], d.get('dp_check'))
PY
