#!/bin/bash
# compute-sanitizer: memcheck over the small-architecture parity cases, the window kernel in both schedules (full architecture,
# nfe = 4) and the audio projection; racecheck (shared-memory hazards) over a small-architecture clip and the window kernel
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_static or small_dynamic or small_rcfg or proj_last or (window_kernel_matches and euler-4)" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?"; tail -6 gpurun_out/sanitize_memcheck.log
timeout -s KILL 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(bf16_mode and small_static) or (window_kernel_matches and euler-3)" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?"; tail -6 gpurun_out/sanitize_racecheck.log
