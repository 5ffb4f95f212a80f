#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q 2>&1 | grep -v "^  \|^$\|^array\|^       " | tail -40 > gpurun_out/m3_tests.log; cat gpurun_out/m3_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/m3_smoke.log 2>&1; tail -5 gpurun_out/m3_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m3_bench.json 2> gpurun_out/m3_bench.err
tail -c 2600 gpurun_out/m3_bench.json; tail -5 gpurun_out/m3_bench.err
