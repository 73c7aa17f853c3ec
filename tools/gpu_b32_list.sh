#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_b32_warm_r2.csv python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline --large-clips 0 > gpurun_out/ncu_b32_warm_r2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_b32_warm_r2.csv | tee gpurun_out/launches_b32_warm_r2.txt | head -14
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 100 > gpurun_out/clocks_b32.csv &
SMI=$!
timeout 600 python bench.py --batch 32 --frames 50 --steps 20 --warmup 5 --no-cpu-baseline --large-clips 0 > gpurun_out/bench_b32_f50.json 2> gpurun_out/bench_b32_f50.err
kill $SMI
python -c "
import json; d=json.loads(open('gpurun_out/bench_b32_f50.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['us_per_ode_step'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])"
sort gpurun_out/clocks_b32.csv | uniq -c | sort -rn | head -5
