#!/bin/bash
# round-end style pass: all GPU tests, smoke, both bench arms, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$\|^array\|^       " | tail -6 > gpurun_out/full_tests.log; cat gpurun_out/full_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/full_smoke.log 2>&1; tail -2 gpurun_out/full_smoke.log
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; tail -c 300 gpurun_out/full_bench.json; tail -3 gpurun_out/full_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/full_bench_ref.json 2> gpurun_out/full_bench_ref.err; tail -c 400 gpurun_out/full_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1e.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
