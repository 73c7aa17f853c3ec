#!/bin/bash
# grouped window kernel: parity, then a sweep of its knobs (lookahead, K splits, slice widths)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "window_kernel_matches" > gpurun_out/pytest_g2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_g2.log; tail -5 gpurun_out/pytest_g2.log
run() {  # label, env...
  local label="$1"; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label: us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']))
"
}
run "split-K (FMT_WINDOW=1)" FMT_WINDOW=1
run "grouped default" FMT_WINDOW=2
run "grouped la=2" FMT_WINDOW=2 FMT_WIN_LA=2
run "grouped la=3" FMT_WINDOW=2 FMT_WIN_LA=3
run "grouped la=1000" FMT_WINDOW=2 FMT_WIN_LA=1000
run "grouped pk=2,2,0,3" FMT_WINDOW=2 FMT_WIN_PK=2,2,0,3
run "grouped pk=2,2,0,3 la=2" FMT_WINDOW=2 FMT_WIN_PK=2,2,0,3 FMT_WIN_LA=2
run "grouped pk=1,1,1,2" FMT_WINDOW=2 FMT_WIN_PK=1,1,1,2
run "grouped pk=1,1,1,4 la=2" FMT_WINDOW=2 FMT_WIN_PK=1,1,1,4 FMT_WIN_LA=2
run "grouped nofuse" FMT_WINDOW=2 FMT_WIN_FUSE_GELU=0
FMT_WINDOW=2 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g2.txt 2>&1; head -75 gpurun_out/win_trace_g2.txt | tail -72
