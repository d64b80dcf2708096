#!/bin/bash
cd /root/repo
timeout 200 python -m pytest tests/test_gpu_conv_gemm.py -x -q -m gpu 2>&1 | tail -8
echo "--- pair+halo (default)"
timeout 120 python tools/bench_gemm.py --only "head 1" 2>&1 | tail -4
timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
echo "--- pair per-tap (PAIR_HALO=0)"
HIS_GEMM_PAIR_HALO=0 timeout 120 python tools/bench_gemm.py --only "head 1" 2>&1 | tail -4
HIS_GEMM_PAIR_HALO=0 timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
echo "--- pair+halo only for N=128 (PAIR_HALO=128 w/ astages 2)"
HIS_GEMM_ASTAGES=2 timeout 120 python tools/bench_gemm.py --only "head 1" 2>&1 | tail -4
HIS_GEMM_ASTAGES=2 timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
