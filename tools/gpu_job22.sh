#!/bin/bash
for d in 0 1 2 8 10 15; do echo "--- DEBUG=$d"; HIS_GEMM_DEBUG=$d timeout 100 python tools/bench_gemm.py --only "head 64->64" --reps 20 2>&1; done
for d in 0 1 2 8 10 15; do echo "--- DEBUG=$d"; HIS_GEMM_DEBUG=$d timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" --reps 20 2>&1; done
