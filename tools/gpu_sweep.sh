#!/bin/bash
for b in 1 2 4 8 16 32 64; do
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --batch $b --frames 100 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('B', d['config']['clips_per_gpu'], 'us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'tensor-frac', round(d['roofline']['frac'] if d['roofline']['bound']=='tensor' else d['roofline']['other_bound_frac'],3), 'graph nodes', d['graph_kernel_nodes'])
"
done
