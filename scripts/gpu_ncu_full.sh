#!/bin/bash
# round 2: ncu --set full captures of the dominant kernels (one gpurun call; every ncu command follows a plain run of the
# same command).  The .ncu-rep files are summarised ON THE BOX (gpurun_out/ is limited to 64 MiB) and removed.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_bwd -s 354 -c 6 -o /tmp/r2_gru_bwd_enc $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gru_mma_fwd2 -s 207 -c 3 -o /tmp/r2_gru_fwd_enc $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:k_gemm_tc2|k_ce_reg|k_adam" -s 700 -c 36 -o /tmp/r2_gemm_ce_adam $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu3 rc=$?"
ECMD="python scripts/gpu_embed_prof.py"
$ECMD > gpurun_out/ncu_embed_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gru_tc_fwd3 -s 9 -c 3 -o /tmp/r2_gru_tc_fwd3 $ECMD > gpurun_out/ncu4.log 2>&1
echo "ncu4 rc=$?"
for n in r2_gru_bwd_enc r2_gru_fwd_enc r2_gemm_ce_adam r2_gru_tc_fwd3; do
  python scripts/summarize_ncu.py full /tmp/$n.ncu-rep gpurun_out/${n}_ncu.md > /dev/null 2>&1
  ncu -i /tmp/$n.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/${n}_source.csv
  ls -la gpurun_out/${n}_ncu.md gpurun_out/${n}_source.csv
done
du -sh gpurun_out
