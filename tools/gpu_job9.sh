#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 8 3 7 15; do echo "--- DEBUG=$d"; HIS_GEMM_DEBUG=$d timeout 100 python tools/bench_gemm.py --only "dec4 conv2" 2>&1; HIS_GEMM_DEBUG=$d timeout 100 python tools/bench_gemm.py --only "head 64->64" 2>&1; done | tee gpurun_out/s5_debug_sweep.log
