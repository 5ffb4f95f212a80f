#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_bf16.py tests/test_gpu_parity.py "tests/test_gpu_configs.py::test_scaled_config_dims_match_oracle" -q -x 2>&1 ) | tail -3
python bench.py --workload scaled --steps 3 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('scaled', d['ms_per_step'], d['phases_ms'])"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C1', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'blocking', round(d['e2e']['blocking']['ms_per_step'],3))"
