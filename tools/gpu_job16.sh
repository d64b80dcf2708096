#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_gemm.py -x -q > gpurun_out/s11_gemm_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s11_gemm_tests.log
tail -5 gpurun_out/s11_gemm_tests.log
grep -q "rc=0" gpurun_out/s11_gemm_tests.log || exit 1
echo "--- halo on"; timeout 120 python tools/exp_cin72.py 2>&1 | head -3
echo "--- halo on"; timeout 200 python tools/bench_gemm.py 2>&1 | tee gpurun_out/s11_sweep_halo1.log
for d in 15; do echo "--- DEBUG=$d"; HIS_GEMM_DEBUG=$d timeout 100 python tools/bench_gemm.py --only "k3" 2>&1; done | tee gpurun_out/s11_debug_sweep.log
