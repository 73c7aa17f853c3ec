#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_batch" > gpurun_out/pytest_lb.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_lb.log

for i in 1 2; do
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --batch 32 --frames 200 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d['roofline']['bound'], round(d['roofline']['frac'],3), d['clocks'])
    else: print(l.rstrip())
"
done
