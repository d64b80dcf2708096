#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "pipelined or graph" > gpurun_out/s21_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s21_tests.log
tail -12 gpurun_out/s21_tests.log
for f in "" "--no-pipeline"; do echo "=== $f"; python bench.py --steps 5 --warmup 3 --no-cpu-baseline $f 2> gpurun_out/s21_b0.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"; tail -2 gpurun_out/s21_b0.err; done
