#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$\|^array\|^       " | tail -40 > gpurun_out/all_tests.log; cat gpurun_out/all_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
tail -c 1800 gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err
