#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s19_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s19_tests.log
tail -12 gpurun_out/s19_tests.log
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline 2> gpurun_out/s19_b0.err | cut -c1-180; head -22 gpurun_out/s19_b0.err; grep -E "cin8 |spatial" gpurun_out/s19_b0.err
