#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gru_tc.py -q -x 2>&1 ) > gpurun_out/r2f_tc_tests.log
echo "tc tests rc=$?" >> gpurun_out/r2f_tc_tests.log
grep -E "^E |passed|failed|rc=" gpurun_out/r2f_tc_tests.log | tail -8
for mode in 1; do
  ARGSIM_GRU_TC=$mode ARGSIM_ENC_BWD_CHUNK=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_tc$mode.json 2> gpurun_out/r2f_bench_tc$mode.err
  echo "bench tc$mode rc=$?"; tail -3 gpurun_out/r2f_bench_tc$mode.err
done
ARGSIM_GRU_TC=1 ARGSIM_GRU_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/r2f_prof_tc.err
grep "gru_tc_prof" gpurun_out/r2f_prof_tc.err | sort | uniq -c | sort -rn | head -8
python - <<'PY'
import json
for f in ('r2f_bench_tc1',):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), {k: v['ms_per_step'] for k, v in d['kernels'].items() if k.startswith('gru')}, 'loss', d['last_step']['loss'])
        print('   embed', round(d['embed']['value']), 'seq/s', round(d['embed']['ms_per_batch'], 2), 'ms;  strong b512', round(d['strong_scaling']['ms_per_step'], 2), 'ms', d['strong_scaling']['phases_ms'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
