#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q > gpurun_out/s25_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s25_tests.log
tail -4 gpurun_out/s25_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/s25_b0.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'blocking', d['e2e']['blocking_value'], d['launches_per_step'])"; tail -3 gpurun_out/s25_b0.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload b1 2> gpurun_out/s25_b1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'blocking', d['e2e']['blocking_value'])"; tail -3 gpurun_out/s25_b1.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload b0_160x120 2> gpurun_out/s25_b0s.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'blocking', d['e2e']['blocking_value'])"; tail -3 gpurun_out/s25_b0s.err
nvidia-smi --query-gpu=memory.used --format=csv
