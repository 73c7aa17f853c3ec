#!/bin/bash
mkdir -p gpurun_out
export FMT_FLOW_SPIN_MS=200
for pm in 0 1; do
FMT_FLOW_POLL=$pm timeout 300 python tools/flow_check.py 1 2>&1 | grep "FMT_WINDOW=3"
FMT_FLOW_POLL=$pm FLOW_TRACE_CTA=17 timeout 200 python tools/flow_trace.py 1 v > gpurun_out/flow_trace_poll$pm.txt 2>&1
grep -A12 "^GEMM engine" gpurun_out/flow_trace_poll$pm.txt | grep -v "^ *[0-9.]* *[GS] "; grep "^chunk\|evaluation span" gpurun_out/flow_trace_poll$pm.txt; grep -A6 "^spread" gpurun_out/flow_trace_poll$pm.txt
done
