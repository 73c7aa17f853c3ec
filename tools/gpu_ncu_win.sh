#!/bin/bash
mkdir -p gpurun_out
# (1) launch list of one B=1 sampler call (window kernel + prepare kernels), (2) full capture of the window kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_b1_window.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
echo "ncu list exit $?"
python tools/launch_summary.py gpurun_out/launches_b1_window.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmt_window_kernel -s 2 -c 1 -o gpurun_out/prof_window -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_win.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full_win.log; ls -la gpurun_out/prof_window.ncu-rep
