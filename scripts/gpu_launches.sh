#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload scaled --steps 3 --warmup 2 > gpurun_out/launches_scaled1.json 2> gpurun_out/launches_scaled1.err
echo "scaled rc=$?"; tail -2 gpurun_out/launches_scaled1.err; cut -c1-700 gpurun_out/launches_scaled1.json
# launch list of the default bench command (every launch, device time), after a plain run of the same command exited 0
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/launches_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/launches_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_launches.csv
