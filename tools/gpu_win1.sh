#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_kernel or bf16_mode or properties" > gpurun_out/pytest_win.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_win.log
timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_win_b1.json 2> gpurun_out/bench_win_b1.err
echo "bench exit $?"; cat gpurun_out/bench_win_b1.json; tail -5 gpurun_out/bench_win_b1.err
