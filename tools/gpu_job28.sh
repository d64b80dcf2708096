#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_gemm.py -x -q > gpurun_out/s20_gemm_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s20_gemm_tests.log
tail -3 gpurun_out/s20_gemm_tests.log
grep -q "rc=0" gpurun_out/s20_gemm_tests.log || exit 1
timeout 200 python tools/bench_gemm.py --only "k3" 2>&1 | tee gpurun_out/s20_sweep.log
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline 2> gpurun_out/s20_b0.err | cut -c1-180; head -4 gpurun_out/s20_b0.err
