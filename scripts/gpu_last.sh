#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/full_bench.json; tail -3 gpurun_out/full_bench.err
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
