#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 60 -c 4 -o gpurun_out/prof_tc2 -f python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo ncu exit $?
ls -la gpurun_out/*.ncu-rep
