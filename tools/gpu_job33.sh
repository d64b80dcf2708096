#!/bin/bash
cd /root/repo
timeout 200 python -m pytest tests/test_gpu_conv_gemm.py -x -q -m gpu 2>&1 | tail -15
echo "--- pair on (default)"
timeout 120 python tools/bench_gemm.py --only "head 128->128" 2>&1 | tail -3
timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
echo "--- pair for N>=256 too"
HIS_GEMM_PAIR=128 timeout 120 python tools/bench_gemm.py --only "head 256->256" 2>&1 | tail -3
echo "--- pair off"
HIS_GEMM_PAIR=0 timeout 120 python tools/bench_gemm.py --only "head 128->128" 2>&1 | tail -3
HIS_GEMM_PAIR=0 timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
nvidia-smi --query-gpu=name,memory.used --format=csv
