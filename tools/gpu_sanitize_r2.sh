#!/bin/bash
# round 2: compute-sanitizer over every schedule of the one-clip regime (dataflow kernel with 64- and 32-row chunks, round 1's two window
# kernels, one kernel per op): memcheck (full architecture, nfe = 4, 3 windows; plus the batched per-op path with 4 clips, the 4-branch
# guidance and one forward_with_cfv evaluation - eight-warp GEMM epilogues, deduplicated condition rows) and racecheck (single branch, nfe = 3)
mkdir -p gpurun_out
export FMT_FLOW_SPIN_MS=120000
timeout -s KILL 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(window_kernel_matches and euler-4) or properties_full_size or small_rcfg or (cfv_step4 and bf16)" > gpurun_out/sanitize_memcheck_r2.log 2>&1
echo "memcheck exit $?"; tail -6 gpurun_out/sanitize_memcheck_r2.log
timeout -s KILL 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_kernel_matches and euler-3" > gpurun_out/sanitize_racecheck_r2.log 2>&1
echo "racecheck exit $?"; tail -6 gpurun_out/sanitize_racecheck_r2.log
