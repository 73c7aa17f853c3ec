#!/bin/bash
# grouped window kernel: which of MMA / activation loads / weight loads paces a GEMM stage (timing only, results invalid)
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
run "grouped no-mma" FMT_WINDOW=2 FMT_WIN_DBG=1
run "grouped no-act-loads" FMT_WINDOW=2 FMT_WIN_DBG=2
run "grouped no-weight-loads" FMT_WINDOW=2 FMT_WIN_DBG=4
run "grouped no loads" FMT_WINDOW=2 FMT_WIN_DBG=6
run "grouped nothing" FMT_WINDOW=2 FMT_WIN_DBG=7
