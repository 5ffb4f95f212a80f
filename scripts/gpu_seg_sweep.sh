#!/bin/bash
# C1 step under other segment lengths of the decoder wavefront / the encoder's BPTT chains
run() { env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['phases_ms']; print('$*', round(d['ms_per_step'],3), 'dec', round(p['dec_fwd'],3), round(p['dec_bwd'],3), 'enc', round(p['enc_fwd'],3), round(p['enc_bwd'],3))"; }
run X=0
run ARGSIM_DEC_SEG=48
run ARGSIM_DEC_SEG=86
run ARGSIM_DEC_SEG=103
run ARGSIM_ENC_SEG=128
run ARGSIM_ENC_SEG=256
run X=1
