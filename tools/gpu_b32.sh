#!/bin/bash
# 32-clip regime: parity of the batched path, launch list of one step, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "large_batch or configs3 or dynamic_emotion or properties" > gpurun_out/pytest_b32.log 2>&1; tail -3 gpurun_out/pytest_b32.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_b32_r2.csv python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline --large-clips 0 > gpurun_out/ncu_b32_r2.log 2>&1
echo ncu exit $?
python tools/launch_summary.py gpurun_out/launches_b32_r2.csv | tee gpurun_out/launches_b32_r2.txt | head -12
timeout 600 python bench.py --batch 32 --frames 200 --steps 5 --warmup 3 --no-cpu-baseline --large-clips 0 > gpurun_out/bench_b32_r2.json 2> gpurun_out/bench_b32_r2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_b32_r2.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['us_per_ode_step'], d['roofline']['frac'], d['clocks'])"
