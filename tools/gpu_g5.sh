#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "window_kernel_matches" > gpurun_out/pytest_g5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_g5.log; tail -4 gpurun_out/pytest_g5.log
run() {
  local label="$1"; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label: us/step', round(d['us_per_ode_step'],1))
"
}
run "grouped" FMT_WINDOW=2
run "grouped la=2" FMT_WINDOW=2 FMT_WIN_LA=2
run "grouped nothing (commits)" FMT_WINDOW=2 FMT_WIN_DBG=7
run "grouped nothing (plain arrives)" FMT_WINDOW=2 FMT_WIN_DBG=15
run "grouped pk=2,2,0,3" FMT_WINDOW=2 FMT_WIN_PK=2,2,0,3
FMT_WINDOW=2 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g5.txt 2>&1
FMT_WINDOW=2 FMT_WIN_DBG=15 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g5_nothing.txt 2>&1
grep -B 14 -A 12 "sum of spans" gpurun_out/win_trace_g5.txt | grep -v Warn
echo ---- nothing; grep -B 10 -A 8 "sum of spans" gpurun_out/win_trace_g5_nothing.txt | grep -v Warn
