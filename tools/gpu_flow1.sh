#!/bin/bash
# first GPU contact of the dataflow kernel: parity against the per-op path, then the suite and the bench
mkdir -p gpurun_out
FMT_FLOW_SPIN_MS=200 timeout 300 python tools/flow_check.py > gpurun_out/flow_check.log 2>&1; echo "flow_check rc=$?" >> gpurun_out/flow_check.log
tail -20 gpurun_out/flow_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -8 gpurun_out/pytest_r2.log
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/bench_r2_ours.json 2> gpurun_out/bench_r2_ours.err; tail -c 2500 gpurun_out/bench_r2_ours.json; tail -5 gpurun_out/bench_r2_ours.err
