#!/bin/bash
echo "== M=5760"; GB_BN=512,256,128 timeout -s KILL 200 python tools/gemm_bench.py
echo "== M=46080"; GB_M=46080 GB_BN=512,256 timeout -s KILL 200 python tools/gemm_bench.py
for pair in 0 1; do echo "== FMT_PAIR=$pair B=32"
FMT_PAIR=$pair timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --batch 32 --frames 200 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d['roofline']['bound'], round(d['roofline']['frac'],3), d['clocks'])
    else: print(l.rstrip())
"
done
