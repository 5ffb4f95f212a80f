#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gru_tc.py -q -x 2>&1 ) > gpurun_out/r2h_tc_tests.log
echo "tc tests rc=$?" >> gpurun_out/r2h_tc_tests.log
grep -E "^E |passed|failed|rc=" gpurun_out/r2h_tc_tests.log | tail -4
for dl in 0; do
  ARGSIM_GRU_TC_DELAY=$dl ARGSIM_GRU_TC=1 ARGSIM_ENC_BWD_CHUNK=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2h_bench_d$dl.json 2> gpurun_out/r2h_bench_d$dl.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2h_bench_d$dl.json').read().strip().splitlines()[-1])
print('delay $dl ms', round(d['ms_per_step'], 3), {k: v['ms_per_step'] for k, v in d['kernels'].items() if k.startswith('gru_fwd')})
PY
done
ARGSIM_GRU_TC_DELAY=0 ARGSIM_GRU_TC=1 ARGSIM_GRU_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/r2h_prof_tc.err
grep "gru_tc_prof" gpurun_out/r2h_prof_tc.err | sort | uniq -c | sort -rn | head -3
