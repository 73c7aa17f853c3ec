#!/bin/bash
mkdir -p gpurun_out
FMT_WINDOW=2 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g4.txt 2>&1
FMT_WINDOW=2 FMT_WIN_DBG=7 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g4_nothing.txt 2>&1
FMT_WINDOW=1 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g4_v1.txt 2>&1
grep -A 30 "sum of spans" gpurun_out/win_trace_g4.txt | grep -v Warn | head -24
echo ---- nothing; grep -B 12 -A 20 "sum of spans" gpurun_out/win_trace_g4_nothing.txt | grep -v Warn | head -40
