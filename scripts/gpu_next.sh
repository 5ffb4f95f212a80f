#!/bin/bash
# First GPU call of the next round: knobs of the side-stream schedule that were set with one measurement each
# (DESIGN.md section 6 / A/B table), swept on one box so that box class does not blur the comparison.  ~15 s per line.
mkdir -p gpurun_out
run() {
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/next_bench.json 2> gpurun_out/next_bench.err || tail -3 gpurun_out/next_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/next_bench.json').read().strip().splitlines()[-1])
k=d['kernels']
print('$1 | ms %.3f | e2e %.3f | enc_bwd %.3f (gru %.3f) dec_bwd %.3f (gru %.3f) logits %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'],
      d['phases_ms']['enc_bwd'], k['gru_bwd_enc']['ms_per_step'], d['phases_ms']['dec_bwd'], k['gru_bwd_dec']['ms_per_step'], d['phases_ms']['logits_ce']))
PY
}
run "default                 "
# one A/B at the end of round 1 (enc BPTT 3.56 -> 3.36 ms, decoder slower): 8-row chunks, a 16-row slice = two alternating one-tile exchanges.
# Run the parity tests with it FIRST; if green, compare and then retry the slice budgets with 16-row slices everywhere.
ARGSIM_GRU_CHUNK=8 timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
ARGSIM_ENC_BWD_CHUNK=8 timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
ARGSIM_ENC_BWD_CHUNK=8 run "chunk 8 in enc BPTT only"
ARGSIM_GRU_CHUNK=8 run "chunk 8                 "
ARGSIM_GRU_CHUNK=8 ARGSIM_NO_SLICE_BUDGET=1 run "chunk 8, 16-row slices  "
ARGSIM_GRU_CHUNK=8 ARGSIM_NO_SLICE_BUDGET=1 ARGSIM_ENC_SEG=0 run "chunk 8, no enc segments"
ARGSIM_GRU_CHUNK=8 ARGSIM_GROUP_CAP=6 run "chunk 8, 6 groups       "
for u in 16 24 48 64; do ARGSIM_SIDE_UNITS=$u run "side units $u           "; done
for e in 128 205 256; do ARGSIM_ENC_SEG=$e run "enc seg $e             "; done
for d in 48 96; do ARGSIM_DEC_SEG=$d run "dec seg $d              "; done
ARGSIM_LOGIT_CHUNK=2048 run "logit chunk 2048        "
ARGSIM_LOGIT_CHUNK=8192 run "logit chunk 8192        "
run "default                 "
