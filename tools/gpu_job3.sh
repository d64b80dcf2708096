#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_tests3.log
tail -3 gpurun_out/s2_tests3.log
python tools/bench_gemm.py --stages 0 > gpurun_out/s2_gemm_sweep3.log 2>&1
cat gpurun_out/s2_gemm_sweep3.log
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/s2_bench_b0.json 2> gpurun_out/s2_bench_b0.err
cut -c1-200 gpurun_out/s2_bench_b0.json
head -70 gpurun_out/s2_bench_b0.err
python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b7 > gpurun_out/s2_bench_b7.json 2> gpurun_out/s2_bench_b7.err
cut -c1-200 gpurun_out/s2_bench_b7.json; head -8 gpurun_out/s2_bench_b7.err
