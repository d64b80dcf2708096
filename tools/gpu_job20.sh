#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q > gpurun_out/s14_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s14_tests.log
tail -6 gpurun_out/s14_tests.log
grep -q "rc=0" gpurun_out/s14_tests.log || exit 1
for t in 0 1; do echo "=== TILED=$t"; HIS_DW_TILED=$t python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline 2> gpurun_out/s14_b0_t$t.err | cut -c1-180; head -3 gpurun_out/s14_b0_t$t.err; grep depthwise gpurun_out/s14_b0_t$t.err | head -20; done
