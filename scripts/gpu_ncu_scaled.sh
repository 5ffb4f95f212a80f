#!/bin/bash
# ncu --set full of the wide-model kernels (BASELINE configs[4] dimensions): the fused per-step forward kernel and the BPTT
# chain's GEMM.  A small batch and a short sequence keep the run short; the kernels' per-launch work is what configs[4] runs
# (64 rows, H = 2048).  The .ncu-rep is summarised on the box.
mkdir -p gpurun_out
CMD="python scripts/gpu_scaled_step.py"
$CMD > gpurun_out/ncu_scaled_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_gru_step_fwd|k_gate_bwd|k_gemm_tc2" -s 150 -c 12 -o /tmp/r2_scaled $CMD > gpurun_out/ncu_scaled.log 2>&1
echo "ncu rc=$?"
python scripts/summarize_ncu.py full /tmp/r2_scaled.ncu-rep gpurun_out/r2_scaled_ncu.md > /dev/null 2>&1
ls -la gpurun_out/r2_scaled_ncu.md; tail -3 gpurun_out/ncu_scaled.log
