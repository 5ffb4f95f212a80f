#!/bin/bash
mkdir -p gpurun_out
( ARGSIM_GRU_TC_FWD=1 timeout 900 python -m pytest tests/test_gpu_gru_tc.py -q -x 2>&1 ) > gpurun_out/r2j_tc_tests.log
echo "tc tests rc=$?" >> gpurun_out/r2j_tc_tests.log
grep -E "^E |passed|failed|rc=" gpurun_out/r2j_tc_tests.log | tail -4
ARGSIM_GRU_TC_FWD=1 ARGSIM_GRU_TC=1 ARGSIM_ENC_BWD_CHUNK=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_v1.json 2> gpurun_out/r2j_bench_v1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2j_bench_v1.json').read().strip().splitlines()[-1])
print('v1 split-issue ms', round(d['ms_per_step'], 3), {k: v['ms_per_step'] for k, v in d['kernels'].items() if k.startswith('gru')})
print('   embed', round(d['embed']['value']), 'seq/s', round(d['embed']['ms_per_batch'], 2), 'ms;  strong b512', round(d['strong_scaling']['ms_per_step'], 2), 'ms', d['strong_scaling']['phases_ms'])
PY
ARGSIM_GRU_TC_FWD=1 ARGSIM_GRU_TC=1 ARGSIM_GRU_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/r2j_prof_tc.err
grep "gru_tc_prof" gpurun_out/r2j_prof_tc.err | sort | uniq -c | sort -rn | head -3
