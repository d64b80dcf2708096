#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/prof_halo.py > gpurun_out/s23_memcheck_halo.log 2>&1; echo "rc=$?" >> gpurun_out/s23_memcheck_halo.log
tail -15 gpurun_out/s23_memcheck_halo.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_post.py tests/test_gpu_kernels.py -x -q > gpurun_out/s23_memcheck_post.log 2>&1; echo "rc=$?" >> gpurun_out/s23_memcheck_post.log
tail -8 gpurun_out/s23_memcheck_post.log
