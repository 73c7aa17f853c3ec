#!/bin/bash
# full GPU check + stage trace of the window kernel
mkdir -p gpurun_out
bash tools/gpu_check.sh
timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_final.txt 2>&1
grep "sum of spans" gpurun_out/win_trace_final.txt; tail -12 gpurun_out/win_trace_final.txt
