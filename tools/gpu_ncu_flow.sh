#!/bin/bash
# round 2: (1) launch list of the B=1 sampler call (dataflow window kernel + prepare kernels), (2) full capture of the dataflow kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_b1_flow.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --large-clips 0 > gpurun_out/ncu_b1_flow.log 2>&1
echo "ncu list exit $?"
python tools/launch_summary.py gpurun_out/launches_b1_flow.csv | tee gpurun_out/launches_b1_flow.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmt_flow_kernel -s 2 -c 1 -o gpurun_out/prof_flow -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --large-clips 0 > gpurun_out/ncu_full_flow.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full_flow.log; ls -la gpurun_out/prof_flow.ncu-rep
