#!/bin/bash
# first GPU pass: fp32 parity tests + fp32 bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/m1_tests.log
cat gpurun_out/m1_tests.log
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m1_bench_fp32.json 2> gpurun_out/m1_bench_fp32.err
tail -c 3000 gpurun_out/m1_bench_fp32.json; tail -5 gpurun_out/m1_bench_fp32.err
