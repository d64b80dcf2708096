#!/bin/bash
mkdir -p gpurun_out
{
timeout 500 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_model.py tests/test_gpu_conv_epilogue.py -m gpu -x -q 2>&1 | tail -3
for v in "HIS_GEMM_DEBUG=32" "HIS_X=0" "HIS_GEMM_DEBUG=32" "HIS_X=0"; do env $v timeout 150 python bench.py --steps 3 --warmup 3 --precision strict --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('strict $v', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['issued_frac'])"; done
} > gpurun_out/exp9.log 2>&1
tail -8 gpurun_out/exp9.log
