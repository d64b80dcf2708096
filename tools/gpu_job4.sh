#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_tests.log
tail -3 gpurun_out/s3_tests.log
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/s3_bench_b0.json 2> gpurun_out/s3_bench_b0.err
cut -c1-200 gpurun_out/s3_bench_b0.json
python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b7 > gpurun_out/s3_bench_b7.json 2> gpurun_out/s3_bench_b7.err
cut -c1-200 gpurun_out/s3_bench_b7.json; head -12 gpurun_out/s3_bench_b7.err
python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b1 > gpurun_out/s3_bench_b1.json 2> gpurun_out/s3_bench_b1.err
cut -c1-200 gpurun_out/s3_bench_b1.json; head -12 gpurun_out/s3_bench_b1.err
python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b0_160x120 > gpurun_out/s3_bench_b0s.json 2> gpurun_out/s3_bench_b0s.err
cut -c1-200 gpurun_out/s3_bench_b0s.json; head -40 gpurun_out/s3_bench_b0s.err
