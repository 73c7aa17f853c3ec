#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('us/step', round(d['us_per_ode_step'],1), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']))
    "; }
run FMT_PDL=1 FMT_SKINNY=1
run FMT_PDL=0 FMT_SKINNY=1
run FMT_PDL=1 FMT_SKINNY=0
run FMT_PDL=0 FMT_SKINNY=0
run FMT_PDL=0 FMT_SKINNY=1 FMT_SK_CLUSTER=1
run FMT_PDL=0 FMT_SKINNY=1 FMT_SK_CLUSTER=2
run FMT_PDL=0 FMT_SKINNY=1 FMT_SK_CLUSTER=4
run FMT_PDL=1 FMT_SKINNY=1 FMT_SK_CLUSTER=1
run FMT_PDL=0 FMT_SKINNY=1 FMT_SK_CLUSTER=1 FMT_SK_CTAS=64
run FMT_PDL=0 FMT_SKINNY=1 FMT_SK_CLUSTER=1 FMT_SK_CTAS=148
FMT_PDL=0 FMT_SKINNY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/launches_sk.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sk.log 2>&1
echo ncu exit $?
