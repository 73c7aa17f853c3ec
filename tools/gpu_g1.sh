#!/bin/bash
# grouped window kernel: parity against the per-op path, then timing (FMT_WINDOW=1 vs 2)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "window_kernel_matches" > gpurun_out/pytest_g1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_g1.log; tail -15 gpurun_out/pytest_g1.log
for w in 1 2; do
FMT_WINDOW=$w timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_g1_w$w.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('FMT_WINDOW=$w us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']))
"
tail -2 gpurun_out/bench_g1_w$w.err
done
FMT_WINDOW=2 FMT_WIN_FUSE_GELU=0 timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('grouped, separate GELU: us/step', round(d['us_per_ode_step'],1))
"
FMT_WINDOW=2 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g1.txt 2>&1; tail -30 gpurun_out/win_trace_g1.txt
