#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/s2_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_tests2.log
tail -3 gpurun_out/s2_tests2.log
python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b7 > gpurun_out/s2_bench_b7.json 2> gpurun_out/s2_bench_b7.err
cut -c1-200 gpurun_out/s2_bench_b7.json; head -8 gpurun_out/s2_bench_b7.err
python tools/prof_set.py > gpurun_out/prof_set_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"depthwise|conv_gemm" -c 15 -o gpurun_out/s2_prof_set python tools/prof_set.py > gpurun_out/prof_set_ncu.log 2>&1
tail -3 gpurun_out/prof_set_ncu.log
ls -la gpurun_out/*.ncu-rep
