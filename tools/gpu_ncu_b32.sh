#!/bin/bash
# round 2: full ncu capture of the dominant kernel of the 32-clip step (the CTA-pair GEMM: qkv / fc1 launches)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 60 -c 4 -o gpurun_out/prof_tc2_r2 -f python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline --large-clips 0 > gpurun_out/ncu_full_b32_r2.log 2>&1
echo ncu exit $?
ls -la gpurun_out/prof_tc2_r2.ncu-rep
