#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_conv_gemm.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/bench_gemm.py --only "head 1" 2>&1 | tail -4
timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
timeout 120 python tools/bench_gemm.py --only "expand" --act 2 2>&1 | tail -2
timeout 120 python tools/bench_gemm.py --only "convT" 2>&1 | tail -2
