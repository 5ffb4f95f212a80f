#!/bin/bash
# round 2, call 1: whole GPU suite (incl. the new full-size / config tests), chunk-8 BPTT parity, bench A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 ) > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -30 gpurun_out/r2a_tests.log
( ARGSIM_ENC_BWD_CHUNK=8 timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py "tests/test_gpu_configs.py::test_c1_full_size_all_gradients_match_torch_autograd" -q -x ) > gpurun_out/r2a_chunk8_tests.log 2>&1
echo "chunk8 tests rc=$?" >> gpurun_out/r2a_chunk8_tests.log
tail -5 gpurun_out/r2a_chunk8_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
ARGSIM_ENC_BWD_CHUNK=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2a_bench_chunk8.json 2> gpurun_out/r2a_bench_chunk8.err
echo "bench chunk8 rc=$?"
python - <<'PY'
import json
for f in ('r2a_bench', 'r2a_bench_chunk8'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), {k: v['ms_per_step'] for k, v in d['kernels'].items() if k.startswith('gru')})
        for k in ('embed', 'strong_scaling', 'cpu_baseline'):
            if k in d:
                print('  ', k, {a: b for a, b in d[k].items() if a in ('value', 'ms_per_step', 'ms_per_batch', 'cores', 'roofline')})
    except Exception as e:
        print(f, 'unreadable', e)
PY
