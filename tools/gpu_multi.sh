#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --batch 32 --frames 200 > gpurun_out/bench_n${N}_b32.json 2> gpurun_out/bench_n${N}_b32.err
echo "exit $?"; cat gpurun_out/bench_n${N}_b32.json; tail -3 gpurun_out/bench_n${N}_b32.err
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_b1.json 2> gpurun_out/bench_n${N}_b1.err
echo "exit $?"; cat gpurun_out/bench_n${N}_b1.json; tail -3 gpurun_out/bench_n${N}_b1.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | tail -2
