#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gru_tc.py tests/test_gpu_bf16.py tests/test_gpu_gemm.py -q -x 2>&1 ) > gpurun_out/r2m_tc_tests.log
echo "tc tests rc=$?" >> gpurun_out/r2m_tc_tests.log
grep -E "^E |passed|failed|rc=" gpurun_out/r2m_tc_tests.log | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'blocking', round(d['e2e']['blocking']['ms_per_step'], 3))
for k, v in sorted(d['kernels'].items()):
    print('  %-28s %8.4f ms  frac %.3f' % (k, v['ms_per_step'], v['frac']))
print('   embed', round(d['embed']['value']), 'seq/s', round(d['embed']['ms_per_batch'], 2), 'ms;  strong b512', round(d['strong_scaling']['ms_per_step'], 2), 'ms', d['strong_scaling']['phases_ms'])
PY
