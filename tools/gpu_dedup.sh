#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log; tail -4 gpurun_out/pytest_r2.log
for dd in 0 1; do
echo "FMT_DEDUP=$dd"
FMT_DEDUP=$dd timeout 300 python tools/flow_check.py 1 2>&1 | grep "FMT_WINDOW=3\|FMT_WINDOW=0"
FMT_DEDUP=$dd timeout 600 python bench.py --batch 32 --frames 200 --steps 5 --warmup 3 --no-cpu-baseline --large-clips 256 > gpurun_out/bench_b32_dd$dd.json 2> gpurun_out/bench_b32_dd$dd.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_b32_dd$dd.json').read().strip().splitlines()[-1]); print('32 clips', d['value'], d['roofline']['us_per_ode_step'], d['roofline']['frac'], d['clocks']['sm_mhz']); l=d['large_batch']; print('256 clips', l['value'], l['us_per_ode_step'], l['roofline']['frac'], l['clocks']['sm_mhz'])"
done
