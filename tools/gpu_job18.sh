#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1b_tests.log
tail -3 gpurun_out/r1b_tests.log
python bench.py --workload post --steps 5 --warmup 3 > gpurun_out/r1b_bench_post.json 2> gpurun_out/r1b_bench_post.err; cut -c1-600 gpurun_out/r1b_bench_post.json; tail -3 gpurun_out/r1b_bench_post.err
python bench.py --steps 5 --warmup 3 --breakdown > gpurun_out/r1b_bench_b0.json 2> gpurun_out/r1b_bench_b0.err; cut -c1-200 gpurun_out/r1b_bench_b0.json
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 639 -c 213 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_launches.log 2>&1
tail -2 gpurun_out/r1b_ncu_launches.log
python tools/prof_set.py > gpurun_out/r1b_prof_set_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"depthwise|conv_gemm|mask_cleanup" -c 16 -o gpurun_out/r1b_prof_set python tools/prof_set.py > gpurun_out/r1b_prof_set_ncu.log 2>&1
tail -3 gpurun_out/r1b_prof_set_ncu.log
ls -la gpurun_out/*.ncu-rep
