#!/bin/bash
# one full step of the K-major x K-major tcgen05 GEMMs (input projections + the three vocab-projection chunks), --set full
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"k_gemm_tc2<.int.0, .int.0>" -s 54 -c 27 -f -o gpurun_out/prof_gemm00_r1d $CMD > gpurun_out/ncu_g.log 2>&1
echo "gemm00 rc=$?"; ls -la gpurun_out/prof_gemm00_r1d.ncu-rep
