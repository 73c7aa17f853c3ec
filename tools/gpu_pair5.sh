#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for sk in 0 1; do echo "== FMT_SPLITK=$sk"
FMT_SPLITK=$sk timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --batch 32 --frames 200 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d['roofline']['bound'], round(d['roofline']['frac'],3), d['clocks'])
    else: print(l.rstrip())
"
done
