#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_gemm.py -x -q > gpurun_out/s24_t1.log 2>&1; echo "rc=$?" >> gpurun_out/s24_t1.log; tail -2 gpurun_out/s24_t1.log
HIS_GEMM_DIRECT=3 timeout 300 python -m pytest tests/test_gpu_conv_gemm.py -x -q > gpurun_out/s24_t3.log 2>&1; echo "rc=$?" >> gpurun_out/s24_t3.log; tail -2 gpurun_out/s24_t3.log
for d in 1 3; do echo "=== DIRECT=$d"; HIS_GEMM_DIRECT=$d python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b1 2> gpurun_out/s24_b1_d$d.err | cut -c1-150; grep -E "cin256 cout256 k3" gpurun_out/s24_b1_d$d.err | head -3; done
