#!/bin/bash
mkdir -p gpurun_out
export FMT_FLOW_SPIN_MS=200
timeout 300 python tools/flow_check.py 2>&1 | grep -v "FMT_WINDOW=0" | tee gpurun_out/flow_check.log | tail -8
FLOW_TRACE_CTA=17 FMT_FLOW_CH=64 timeout 200 python tools/flow_trace.py 1 v > gpurun_out/flow_trace_ch64.txt 2>&1
grep -A12 "^GEMM engine" gpurun_out/flow_trace_ch64.txt; grep "^chunk\|evaluation span" gpurun_out/flow_trace_ch64.txt
