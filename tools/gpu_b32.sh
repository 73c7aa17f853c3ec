#!/bin/bash
# 32-clip regime: parity of the batched path, GEMM kernels in isolation, launch list of one step (warm caches), bench lines (32 and 256 clips)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "large_batch or configs3 or dynamic_emotion or properties or kernels or projection or gemm" > gpurun_out/pytest_b32.log 2>&1; tail -3 gpurun_out/pytest_b32.log
timeout 300 python tools/gemm_bench.py 5760 > gpurun_out/gemm_bench_m5760_r2.txt 2>&1; cat gpurun_out/gemm_bench_m5760_r2.txt | tail -18
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_b32_warm_r2.csv python bench.py --steps 1 --warmup 3 --batch 32 --frames 50 --no-cpu-baseline --large-clips 0 > gpurun_out/ncu_b32_warm_r2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_b32_warm_r2.csv | tee gpurun_out/launches_b32_warm_r2.txt | head -10
timeout 600 python bench.py --batch 32 --frames 200 --steps 5 --warmup 3 --no-cpu-baseline --large-clips 256 > gpurun_out/bench_b32_r2.json 2> gpurun_out/bench_b32_r2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_b32_r2.json').read().strip().splitlines()[-1]); print('32 clips', d['value'], d['roofline']['us_per_ode_step'], d['roofline']['frac'], d['clocks']); l=d['large_batch']; print('256 clips', l['value'], l['us_per_ode_step'], l['roofline']['frac'], l['clocks'])"
