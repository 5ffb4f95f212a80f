#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
PREC=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check_fp32_$N.log 2>&1; tail -12 gpurun_out/dp_check_fp32_$N.log
PREC=bf16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_check.py > gpurun_out/dp_check_bf16_$N.log 2>&1; tail -12 gpurun_out/dp_check_bf16_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_dp$N.json').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
    print(d['phases_ms'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/bench_dp$N.err
