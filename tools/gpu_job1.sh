#!/bin/bash
# round-1 session-2 job: parity, GEMM ring-depth sweep, B0 bench + breakdown
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_tests.log
tail -5 gpurun_out/s2_tests.log
python tools/bench_gemm.py --stages 4,12,0 > gpurun_out/s2_gemm_sweep.log 2>&1
cat gpurun_out/s2_gemm_sweep.log
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/s2_bench_b0.json 2> gpurun_out/s2_bench_b0.err
cat gpurun_out/s2_bench_b0.json | cut -c1-400
head -30 gpurun_out/s2_bench_b0.err
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline --workload b7 > gpurun_out/s2_bench_b7.json 2> gpurun_out/s2_bench_b7.err
cut -c1-300 gpurun_out/s2_bench_b7.json; head -12 gpurun_out/s2_bench_b7.err
