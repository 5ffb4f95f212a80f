#!/bin/bash
# data parallel step: overlapped gradient buckets (default) against ONE all-reduce behind the backward pass
mkdir -p gpurun_out
N=${1:-8}
P=29530
for v in "" "ARGSIM_DP_ONE_ALLREDUCE=1"; do
  P=$((P+1))
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 4 --no-extra --no-cpu-baseline > gpurun_out/dp_one.json 2> gpurun_out/dp_one.err || tail -3 gpurun_out/dp_one.err
  python - "$v" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/dp_one.json').read().strip().splitlines()[-1])
print('[%s]' % sys.argv[1], 'N', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'dp_check', d.get('dp_check'), d['phases_ms'])
PY
done
