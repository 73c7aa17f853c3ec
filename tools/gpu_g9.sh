#!/bin/bash
mkdir -p gpurun_out
run() {
  local label="$1"; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label: us/step', round(d['us_per_ode_step'],1))
"
}
run "split-K" FMT_WINDOW=1
run "per-op" FMT_WINDOW=0
run "grouped pk=2,2,0,3" FMT_WINDOW=2 FMT_WIN_PK=2,2,0,3
run "grouped pk=2,2,0,4" FMT_WINDOW=2 FMT_WIN_PK=2,2,0,4
run "grouped pk=2,3,0,4" FMT_WINDOW=2 FMT_WIN_PK=2,3,0,4
run "grouped pk=2,2,0,3 la=2" FMT_WINDOW=2 FMT_WIN_PK=2,2,0,3 FMT_WIN_LA=2
run "grouped pk=2,2,2,3 nofuse" FMT_WINDOW=2 FMT_WIN_PK=2,2,2,3 FMT_WIN_FUSE_GELU=0
run "grouped pk=3,3,0,4" FMT_WINDOW=2 FMT_WIN_PK=3,3,0,4
FMT_WINDOW=2 FMT_WIN_PK=2,2,0,3 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_g9.txt 2>&1
sed -n 1,20p gpurun_out/win_trace_g9.txt
grep -A 8 "sum of spans" gpurun_out/win_trace_g9.txt | grep -v Warn
tail -12 gpurun_out/win_trace_g9.txt
