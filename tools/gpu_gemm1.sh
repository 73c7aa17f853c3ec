#!/bin/bash
for gm in 1 4 8 23; do echo "== raster_gm=$gm"; FMT_RASTER_GM=$gm GB_BN=512 timeout -s KILL 200 python tools/gemm_bench.py 2>&1 | grep -v cuBLAS; done
echo "== reference kernels"; GB_BN=256,128 timeout -s KILL 200 python tools/gemm_bench.py
