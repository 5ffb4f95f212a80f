#!/bin/bash
# full round-end style pass: all GPU tests, smoke, both bench arms, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$\|^array\|^       " | tail -25 > gpurun_out/full_tests.log; cat gpurun_out/full_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/full_smoke.log 2>&1; tail -3 gpurun_out/full_smoke.log
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; tail -c 600 gpurun_out/full_bench.json; tail -3 gpurun_out/full_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/full_bench_ref.json 2> gpurun_out/full_bench_ref.err; tail -c 700 gpurun_out/full_bench_ref.json
timeout 600 python bench.py --workload embed --steps 5 --warmup 3 > gpurun_out/full_bench_embed.json 2> gpurun_out/full_bench_embed.err; tail -c 900 gpurun_out/full_bench_embed.json; tail -3 gpurun_out/full_bench_embed.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma -s 36 -c 12 -f -o gpurun_out/prof_gru_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "gru full rc=$?"
$CMD > gpurun_out/ncu_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_ce|k_gemm_tc|k_adam" -s 300 -c 40 -f -o gpurun_out/prof_gemm_ce_r1 $CMD > gpurun_out/ncu_full2.log 2>&1
echo "gemm/ce full rc=$?"
ls -la gpurun_out | tail -12
