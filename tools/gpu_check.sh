#!/bin/bash
# Round check on the B200 box: GPU tests, smoke, then the two bench regimes. Everything under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-test_}" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 400 python bench.py > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
echo "bench b1 exit $?"; cat gpurun_out/bench_b1.json; tail -3 gpurun_out/bench_b1.err
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --batch 32 --frames 200 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err
echo "bench b32 exit $?"; cat gpurun_out/bench_b32.json; tail -3 gpurun_out/bench_b32.err
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -1
