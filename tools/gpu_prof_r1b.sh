#!/bin/bash
# Profile refresh after the no-local-memory / control-code changes: launch list + full capture of the window kernel, stage trace,
# default bench (configs[1]) with the CPU baseline, 32-clip bench, reference arm.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_b1_window_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1_v2.log 2>&1
echo "ncu list exit $?"
python tools/launch_summary.py gpurun_out/launches_b1_window_v2.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmt_window_kernel -s 2 -c 1 -o gpurun_out/prof_window_v2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_win_v2.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full_win_v2.log; ls -la gpurun_out/prof_window_v2.ncu-rep
FMT_WINDOW=1 timeout -s KILL 300 python tools/win_trace.py 1 > gpurun_out/win_trace_r1b.txt 2>&1
timeout -s KILL 400 python bench.py > gpurun_out/bench_b1_r1b.json 2> gpurun_out/bench_b1_r1b.err
echo "bench b1 exit $?"; cat gpurun_out/bench_b1_r1b.json
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --batch 32 --frames 200 --no-cpu-baseline > gpurun_out/bench_b32_r1b.json 2> gpurun_out/bench_b32_r1b.err
echo "bench b32 exit $?"; cat gpurun_out/bench_b32_r1b.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -1
