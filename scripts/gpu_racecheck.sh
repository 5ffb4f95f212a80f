#!/bin/bash
# compute-sanitizer over the smoke step ("race detection").  The recurrence kernels exchange data through
# tagged volatile words in L2 (inter-CTA, by design outside racecheck's shared-memory scope); what racecheck covers here is
# the shared-memory traffic INSIDE a CTA: partial-sum tiles, gate-exchange buffers, operand tiles written by generic stores
# and read by the tensor core (async proxy).  One tool per gpurun call (B200_PROFILING.md); run after a plain smoke passed.
#   gpurun --timeout 900 -- 'bash scripts/gpu_racecheck.sh racecheck'      (or memcheck / synccheck / initcheck)
TOOL=${1:-racecheck}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain smoke failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
timeout 800 compute-sanitizer --tool $TOOL --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "sanitizer rc=$?"; tail -15 gpurun_out/sanitizer_$TOOL.log
