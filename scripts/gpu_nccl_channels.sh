#!/bin/bash
# data parallel step under different NCCL channel caps (NCCL's CTAs share the SMs with the recurrence grids)
mkdir -p gpurun_out
N=${1:-8}
P=29520
for ch in default 2 4 8; do
  P=$((P+1))
  if [ "$ch" = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/nccl_ch.json 2> gpurun_out/nccl_ch.err || tail -3 gpurun_out/nccl_ch.err
  python - "$ch" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/nccl_ch.json').read().strip().splitlines()[-1])
print('channels', sys.argv[1], 'N', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'dp_check', d.get('dp_check'), d['phases_ms'])
PY
done
