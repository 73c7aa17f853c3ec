#!/bin/bash
mkdir -p gpurun_out
export FMT_FLOW_SPIN_MS=500
timeout 300 python tools/flow_check.py 1 2>&1 | grep -v "FMT_WINDOW=0" | tee gpurun_out/flow_check.log | tail -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -6 gpurun_out/pytest_r2.log
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/bench_r2_ours.json 2> gpurun_out/bench_r2_ours.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r2_ours.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['us_per_ode_step'], d['roofline']['frac'], d['e2e'])"; tail -3 gpurun_out/bench_r2_ours.err
