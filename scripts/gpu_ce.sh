#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ce.py tests/test_gpu_bf16.py -m gpu -q -x 2>&1 | tail -15
timeout 300 python scripts/gpu_kern.py 2>&1 | grep "softmax" | cut -c1-200
ARGSIM_CE_SMEM=1 timeout 300 python scripts/gpu_kern.py 2>&1 | grep softmax | cut -c1-200
