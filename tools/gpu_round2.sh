#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/psnr_dump.py
GB_BN=512,256,128 timeout -s KILL 200 python tools/gemm_bench.py > gpurun_out/gemm_bench_m5760.txt 2>&1; cat gpurun_out/gemm_bench_m5760.txt
GB_M=46080 GB_BN=512,256 timeout -s KILL 200 python tools/gemm_bench.py > gpurun_out/gemm_bench_m46080.txt 2>&1; cat gpurun_out/gemm_bench_m46080.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 60 -c 4 -o gpurun_out/prof_tc2_v2 -f python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo ncu exit $?
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_b32_v2.csv python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline > gpurun_out/ncu_b32.log 2>&1
python tools/launch_summary.py gpurun_out/launches_b32_v2.csv
FMT_WIN_LA=1000 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_final.txt 2>&1; grep -E "mean span|sum of" gpurun_out/win_trace_final.txt
