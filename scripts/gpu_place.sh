#!/bin/bash
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/smid_probe scripts/smid_probe.cu && /tmp/smid_probe | cut -c1-400
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from argsim_b200 import _lib
for method in (0, 32, 96):
    for groups in (3, 8):
        cyc, _ = _lib.bench_exchange(method, groups, 4, 4000)
        print('method', method, 'groups', groups, 'rows 4 cycles/round %.0f' % cyc, flush=True)
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('gru')})"
