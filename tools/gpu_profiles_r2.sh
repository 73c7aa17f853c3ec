#!/bin/bash
# refresh of the round-2 profiles after the last kernel changes (one B200)
mkdir -p gpurun_out
bash tools/gpu_ncu_flow.sh > gpurun_out/gpu_ncu_flow.out 2>&1; tail -3 gpurun_out/gpu_ncu_flow.out
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_b32_warm_r2.csv python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline --large-clips 0 > gpurun_out/ncu_b32_warm_r2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_b32_warm_r2.csv | tee gpurun_out/launches_b32_warm_r2.txt | head -8
bash tools/gpu_final_r2.sh 2>&1 | tail -5
