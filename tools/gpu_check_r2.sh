#!/bin/bash
# round-2 acceptance run on one B200: smoke(), the whole -m gpu suite, both bench arms as the driver invokes them
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2.log; tail -2 gpurun_out/smoke_r2.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log; tail -4 gpurun_out/pytest_r2.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 400 gpurun_out/bench_r2_ref.json; echo
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2_ours.json 2> gpurun_out/bench_r2_ours.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r2_ours.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','us_per_ode_step')}, d['roofline']['frac'], d['e2e']['value'], d['clocks']); l=d['large_batch']; print(l['value'], l['roofline']['frac'], l['clocks']); print(d.get('handoff'))"
