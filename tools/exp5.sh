#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
for v in "HIS_GEMM_DEBUG=64" "HIS_X=0" "HIS_GEMM_DEBUG=64" "HIS_X=0"; do env $v timeout 120 python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'])"; done
for w in b7; do for v in "HIS_GEMM_DEBUG=96" "HIS_X=0"; do env $v timeout 120 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w $v', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'])"; done; done
} > gpurun_out/exp5.log 2>&1
tail -30 gpurun_out/exp5.log
