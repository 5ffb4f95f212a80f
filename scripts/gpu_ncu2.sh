#!/bin/bash
# targeted ncu --set full captures (one launch each) of the launches the roofline numbers are about
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
# encoder layer launches: first 3 fwd2 launches / last 3 bwd launches of every step (30 recurrence launches per step)
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_fwd2 -s 60 -c 1 -f -o gpurun_out/prof_gru_fwd_enc_r1c $CMD > gpurun_out/ncu_c1.log 2>&1; echo "fwd enc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_bwd -s 87 -c 1 -f -o gpurun_out/prof_gru_bwd_enc_r1c $CMD > gpurun_out/ncu_c2.log 2>&1; echo "bwd enc rc=$?"
# one full step of K-major x K-major GEMMs: input projections + the three vocab-projection chunks
ncu --set full --clock-control none -k regex:"k_gemm_tc2<0, 0>" -s 50 -c 26 -f -o gpurun_out/prof_gemm00_r1c $CMD > gpurun_out/ncu_c3.log 2>&1; echo "gemm00 rc=$?"
ls -la gpurun_out/*.ncu-rep
