#!/bin/bash
mkdir -p gpurun_out
N=2
for cfg in "ARGSIM_GROUP_CAP=9" "ARGSIM_GROUP_CAP=9 NCCL_MAX_NCHANNELS=4" "ARGSIM_GROUP_CAP=9 NCCL_MAX_NCHANNELS=2" "NCCL_MAX_NCHANNELS=4"; do
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/r2q.json 2> gpurun_out/r2q.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2q.json').read().strip().splitlines()[-1])
print('$cfg', 'ms', round(d['ms_per_step'], 3), 'dp_check', d.get('dp_check'), {k: round(v, 3) for k, v in d['phases_ms'].items() if k in ('dec_bwd', 'enc_bwd', 'allreduce_wait')})
PY
done
