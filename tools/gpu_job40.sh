#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/bench_gemm.py --only "convT" 2>&1 | tail -1
HIS_GEMM_PAIR_T=0 timeout 120 python tools/bench_gemm.py --only "convT" 2>&1 | tail -1
timeout 280 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
HIS_GEMM_PAIR_T=0 timeout 280 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
