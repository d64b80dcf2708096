#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s9_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_tests.log
tail -5 gpurun_out/s9_tests.log
grep -q "rc=0" gpurun_out/s9_tests.log || exit 1
echo "--- halo on"; timeout 120 python tools/exp_cin72.py 2>&1 | tee gpurun_out/s9_exp_halo1.log
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/s9_bench_b0.json 2> gpurun_out/s9_bench_b0.err
cut -c1-200 gpurun_out/s9_bench_b0.json; head -24 gpurun_out/s9_bench_b0.err
python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b7 > gpurun_out/s9_bench_b7.json 2> gpurun_out/s9_bench_b7.err
cut -c1-200 gpurun_out/s9_bench_b7.json; head -8 gpurun_out/s9_bench_b7.err
python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b1 > gpurun_out/s9_bench_b1.json 2> gpurun_out/s9_bench_b1.err
cut -c1-200 gpurun_out/s9_bench_b1.json; head -30 gpurun_out/s9_bench_b1.err
