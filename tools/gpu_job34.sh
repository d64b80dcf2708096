#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "--- pair on (default)"
timeout 120 python tools/bench_gemm.py --only "head 1" 2>&1 | tail -4
timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
echo "--- pair off"
HIS_GEMM_PAIR=0 timeout 120 python tools/bench_gemm.py --only "head 1" 2>&1 | tail -4
HIS_GEMM_PAIR=0 timeout 120 python tools/bench_gemm.py --only "head 256->" 2>&1 | tail -5
timeout 280 python bench.py --steps 5 --warmup 3 --breakdown --top 30 --no-cpu-baseline > gpurun_out/b0_pair.log 2>&1
tail -1 gpurun_out/b0_pair.log | cut -c1-300
HIS_GEMM_PAIR=0 timeout 280 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
