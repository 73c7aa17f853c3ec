#!/bin/bash
mkdir -p gpurun_out
for pk in "0,0,0,0" "4,8,4,8" "4,8,4,16" "3,4,4,8" "6,8,4,12"; do
  echo "=== FMT_WIN_PK=$pk"
  FMT_WIN_LA=1000 FMT_WIN_PK=$pk timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_pk.txt 2>&1; grep -E "marks b3|mean span|sum of" gpurun_out/win_trace_pk.txt
done
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_kernel or bf16_mode or properties" > gpurun_out/pytest_win.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_win.log
