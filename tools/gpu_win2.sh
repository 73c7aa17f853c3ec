#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace.txt 2>&1; tail -16 gpurun_out/win_trace.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_kernel or bf16_mode or properties" > gpurun_out/pytest_win.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_win.log
timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']))
    else: print(l.rstrip())
"
