#!/bin/bash
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_gpu_conv_epilogue.py tests/test_gpu_conv_gemm.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
for v in "HIS_GEMM_DEBUG=96" "HIS_X=0"; do env $v timeout 150 python bench.py --steps 3 --warmup 3 --workload b0_ln --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('b0_ln $v', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'])"; done
for v in "HIS_X=0" "HIS_X=0"; do env $v timeout 120 python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'])"; done
} > gpurun_out/exp7.log 2>&1
tail -12 gpurun_out/exp7.log
