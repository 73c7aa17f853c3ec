#!/bin/bash
mkdir -p gpurun_out
for la in 1 2 1000; do
  echo "=== FMT_WIN_LA=$la"
  FMT_WIN_LA=$la timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_la$la.txt 2>&1; grep -E "marks b3|mean span|sum of" gpurun_out/win_trace_la$la.txt
done
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_kernel or bf16_mode or properties" > gpurun_out/pytest_win.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_win.log
