#!/bin/bash
mkdir -p gpurun_out
FMT_SKINNY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/launches_b32.csv python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline > gpurun_out/ncu_b32.log 2>&1
echo ncu exit $?
FMT_SKINNY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/launches_b1_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1_tc.log 2>&1
echo ncu exit $?
