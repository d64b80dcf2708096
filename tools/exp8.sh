#!/bin/bash
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
for v in "HIS_X=0"; do env $v timeout 150 python bench.py --steps 3 --warmup 3 --workload b0_ln --no-cpu-baseline --breakdown --top 12 2>gpurun_out/exp8_ln_breakdown.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('b0_ln $v', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'], d['time_share_by_op'])"; done
sed -n 3,30p gpurun_out/exp8_ln_breakdown.txt
} > gpurun_out/exp8.log 2>&1
tail -40 gpurun_out/exp8.log
