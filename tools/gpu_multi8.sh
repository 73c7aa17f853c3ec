#!/bin/bash
# 8-GPU scaling check: configs[1] (1 clip per GPU) and configs[3] (32 clips x 8 s per GPU = 256 clips on the box)
mkdir -p gpurun_out
N=${NGPU:-8}
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_b1.json 2> gpurun_out/bench_n${N}_b1.err
echo "n$N b1 exit $?"; tail -1 gpurun_out/bench_n${N}_b1.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], 'GPUs', round(d['value']), 'frames/s', round(d['us_per_ode_step'],1), 'us/step e2e', round(d['e2e']['value']))"
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --batch 32 --frames 200 > gpurun_out/bench_n${N}_b32.json 2> gpurun_out/bench_n${N}_b32.err
echo "n$N b32 exit $?"; tail -1 gpurun_out/bench_n${N}_b32.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], 'GPUs', round(d['value']), 'frames/s', round(d['us_per_ode_step'],1), 'us/step e2e', round(d['e2e']['value']), d['roofline']['frac'])"
tail -3 gpurun_out/bench_n${N}_b32.err
