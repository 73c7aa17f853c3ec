#!/bin/bash
# round-2 sanity: GPU tests, both bench arms (driver-style invocations)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -5 gpurun_out/pytest_r2.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 600 gpurun_out/bench_r2_ref.json
python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/bench_r2_ours.json 2> gpurun_out/bench_r2_ours.err; tail -c 3000 gpurun_out/bench_r2_ours.json; tail -5 gpurun_out/bench_r2_ours.err
