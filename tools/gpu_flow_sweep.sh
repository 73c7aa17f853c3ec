#!/bin/bash
# env sweep of the dataflow kernel's K splits and ring depths (configs[1]); one line per variant
mkdir -p gpurun_out
run() { echo -n "$1: "; env $1 timeout 120 python tools/flow_check.py 1 2>&1 | grep "FMT_WINDOW=3" | sed 's/.*status/status/'; }
run "FMT_FLOW_NA=2"
run "FMT_FLOW_NA=3"
run "FMT_FLOW_NW=6"
run "FMT_FLOW_NW=5"
run "FMT_WIN_PK=5,16,4,16"
run "FMT_WIN_PK=6,8,4,16"
run "FMT_WIN_PK=4,16,4,18"
run "FMT_WIN_PK=6,16,2,18"
run "FMT_WIN_PK=6,16,4,9"
run "FMT_WIN_PK=3,8,2,9"
run "FMT_FLOW_POLL=0"
run "FMT_FLOW_FIXED=0"
