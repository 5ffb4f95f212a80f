#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_dp$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['phases_ms'])
PY
