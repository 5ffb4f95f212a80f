#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + full captures of the top kernels (small -c: the
# gpurun_out merge is capped at 64 MiB)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_fwd -s 9 -c 1 -f -o gpurun_out/prof_gru_fwd_r1 $CMD > gpurun_out/ncu_full_fwd.log 2>&1
echo "gru fwd rc=$?"
$CMD > gpurun_out/ncu_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_bwd -s 60 -c 1 -f -o gpurun_out/prof_gru_bwd_r1 $CMD > gpurun_out/ncu_full_bwd.log 2>&1
echo "gru bwd rc=$?"
$CMD > gpurun_out/ncu_plain4.log 2>&1 && \
ncu --set full --clock-control none -k regex:"k_ce|k_gemm_tc|k_adam|k_embed" -s 420 -c 14 -f -o gpurun_out/prof_gemm_ce_r1 $CMD > gpurun_out/ncu_full2.log 2>&1
echo "gemm/ce full rc=$?"
ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
