#!/bin/bash
# data-parallel check of the side-stream schedule: N-GPU == 1-GPU gradients (bf16 mode), then bench variants
N=${1:-2}
mkdir -p gpurun_out
PREC=bf16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_check.py > gpurun_out/dp_check_bf16_$N.log 2>&1; tail -3 gpurun_out/dp_check_bf16_$N.log
run() {
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_dp$N.json').read().strip().splitlines()[-1])
    print('$1 N', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'blocking', round(d['e2e']['blocking']['value'],1), d['phases_ms'], d['last_step'])
except Exception as e:
    print('bench parse failed', e)
PY
}
run "default        "
ARGSIM_NO_EARLY_ADAM=1 run "no early adam  "
ARGSIM_WGRAD_OVERLAP=0 run "all in line    "
tail -3 gpurun_out/bench_dp$N.err
