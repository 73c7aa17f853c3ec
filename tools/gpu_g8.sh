#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g8.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_g8.log; tail -4 gpurun_out/pytest_g8.log
run() {
  local label="$1"; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label: us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']))
"
}
run "split-K" FMT_WINDOW=1
run "grouped" FMT_WINDOW=2
run "grouped pk=2,2,0,3" FMT_WINDOW=2 FMT_WIN_PK=2,2,0,3
FMT_WINDOW=1 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g8_v1.txt 2>&1
grep -A 12 "sum of spans" gpurun_out/win_trace_g8_v1.txt | grep -v Warn
tail -12 gpurun_out/win_trace_g8_v1.txt
