#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s26_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s26_tests.log
tail -12 gpurun_out/s26_tests.log
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline 2> gpurun_out/s26_b0.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['launches_per_step'])"; head -12 gpurun_out/s26_b0.err
