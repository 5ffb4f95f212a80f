#!/bin/bash
# ncu evidence for profiles/ (round 1, second pass): launch list of a short bench run + full captures of the top kernels.
# Each ncu pass runs only after the same command exited 0 without ncu.  The gpurun_out merge is capped at 64 MiB.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 900 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_fwd2 -s 9 -c 1 -f -o gpurun_out/prof_gru_fwd_r1b $CMD > gpurun_out/ncu_full_fwd.log 2>&1
echo "gru fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_bwd -s 60 -c 1 -f -o gpurun_out/prof_gru_bwd_r1b $CMD > gpurun_out/ncu_full_bwd.log 2>&1
echo "gru bwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_ce_reg|k_adam" -s 12 -c 4 -f -o gpurun_out/prof_ce_adam_r1b $CMD > gpurun_out/ncu_full_ce.log 2>&1
echo "ce/adam rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 360 -c 24 -f -o gpurun_out/prof_gemm_r1b $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
