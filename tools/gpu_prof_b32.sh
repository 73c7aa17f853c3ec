#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_b32_pair.csv python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline > gpurun_out/ncu_b32.log 2>&1
echo ncu exit $?
python tools/launch_summary.py gpurun_out/launches_b32_pair.csv
