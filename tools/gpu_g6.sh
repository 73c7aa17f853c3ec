#!/bin/bash
mkdir -p gpurun_out
FMT_WINDOW=2 FMT_WIN_DBG=16 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g6.txt 2>&1
FMT_WINDOW=2 FMT_WIN_DBG=31 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g6_nothing.txt 2>&1
grep -A 10 "sum of spans" gpurun_out/win_trace_g6.txt | grep -v Warn
echo ---- nothing; grep -A 10 "sum of spans" gpurun_out/win_trace_g6_nothing.txt | grep -v Warn
