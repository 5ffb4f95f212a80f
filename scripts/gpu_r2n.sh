#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gru_tc.py tests/test_gpu_bf16.py "tests/test_gpu_configs.py::test_embed_batch_4096_ibm_shaped_matches_oracle" tests/test_gpu_fullsize.py -q -x 2>&1 ) > gpurun_out/r2n_tests.log
echo "tests rc=$?" >> gpurun_out/r2n_tests.log
grep -E "^E |passed|failed|rc=" gpurun_out/r2n_tests.log | tail -6
for cn in 32 64; do echo "== fwd3 CN=$cn"; ARGSIM_GRU_TC_FWD3_CN=$cn python scripts/gpu_embed_prof.py 2>&1 | tail -5 | cut -c1-250; done
ARGSIM_GRU_TC_FWD3_CN=32 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench.json 2>gpurun_out/r2n_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'], 3), '   embed', round(d['embed']['value']), 'seq/s', round(d['embed']['ms_per_batch'], 2), 'ms;  strong b512', round(d['strong_scaling']['ms_per_step'], 2), 'ms', d['strong_scaling']['phases_ms'])
PY
