#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post.py -x -q > gpurun_out/s13_post_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13_post_tests.log
tail -4 gpurun_out/s13_post_tests.log
python bench.py --workload post --steps 5 --warmup 3 > gpurun_out/s13_bench_post.json 2> gpurun_out/s13_bench_post.err; python -c "
import json; d=json.loads(open('gpurun_out/s13_bench_post.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['avg_launch_ms'], d['e2e']['value'])"
