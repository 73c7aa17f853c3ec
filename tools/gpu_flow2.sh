#!/bin/bash
mkdir -p gpurun_out
export FMT_FLOW_SPIN_MS=200
for df in 0 7 0 7; do
echo "FMT_FLOW_DEFER=$df"
FMT_FLOW_DEFER=$df timeout 300 python tools/flow_check.py 1 2>&1 | grep "FMT_WINDOW=3"
done
FMT_FLOW_DEFER=7 timeout 300 python tools/flow_check.py 2>&1 | grep "FMT_WINDOW=3"
