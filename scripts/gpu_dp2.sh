#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2p_dp$N.json 2> gpurun_out/r2p_dp$N.err
echo "rc=$?"; tail -3 gpurun_out/r2p_dp$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r2p_dp$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'], 3), 'dp_check', d.get('dp_check'))
print(d['phases_ms'])
print('strong', d['strong_scaling']['ms_per_step'], d['strong_scaling']['value'], d['strong_scaling']['phases_ms'])
print('embed', d['embed']['value'])
PY
