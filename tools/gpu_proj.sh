#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "audio_projection" > gpurun_out/pytest_proj.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_proj.log; tail -15 gpurun_out/pytest_proj.log
timeout -s KILL 300 python tools/proj_bench.py 2>&1 | tee gpurun_out/proj_bench.txt | tail -6
