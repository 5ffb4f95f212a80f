#!/bin/bash
# scaled config (BASELINE configs[4]) on one GPU: A/B of an environment switch given as $1 (e.g. ARGSIM_GENERIC_NO_PDL=1)
mkdir -p gpurun_out
for v in "" "$1"; do
  env $v timeout 300 python bench.py --workload scaled --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/scaled_ab.json 2> gpurun_out/scaled_ab.err || tail -3 gpurun_out/scaled_ab.err
  python - "$v" <<'PY'
import json, sys
d = json.loads(open('gpurun_out/scaled_ab.json').read().strip().splitlines()[-1])
print('[%s]' % sys.argv[1], 'ms', round(d['ms_per_step'], 2), d['phases_ms'])
PY
done
