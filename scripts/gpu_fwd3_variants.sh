#!/bin/bash
mkdir -p gpurun_out
for v in 48 6416; do
  ( ARGSIM_GRU_TC_FWD3_CN=$v timeout 600 python -m pytest tests/test_gpu_gru_tc.py -q -x -k "mode4 or 4-" 2>&1 ) | tail -2
done
for v in 64 48 6416; do echo "== fwd3 variant $v"; ARGSIM_GRU_TC_FWD3_CN=$v python scripts/gpu_embed_prof.py 2>&1 | sed -n 2,3p | cut -c1-230; done
