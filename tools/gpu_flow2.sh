#!/bin/bash
mkdir -p gpurun_out
export FMT_FLOW_SPIN_MS=300
timeout 300 python tools/flow_check.py 2>&1 | grep "FMT_WINDOW=3\|FMT_WINDOW=1" | tee gpurun_out/flow_check.log
FLOW_TRACE_TAG=d timeout 200 python tools/flow_trace.py 1 > gpurun_out/flow_trace_d.txt 2>&1
grep -A8 "^GEMM engine" gpurun_out/flow_trace_d.txt; grep "^chunk\|evaluation span" gpurun_out/flow_trace_d.txt
