#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q 2>&1 | grep -v "^  \|^$\|^array\|^       " | tail -40 > gpurun_out/m2_gemm.log; cat gpurun_out/m2_gemm.log
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q 2>&1 | grep -E "AssertionError|passed|failed|Error" | tail -20 > gpurun_out/m2_bf16.log; cat gpurun_out/m2_bf16.log
