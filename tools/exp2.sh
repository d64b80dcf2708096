#!/bin/bash
# fast chunk body + early residual issue: parity first, then A/B against the general body (HIS_GEMM_DEBUG=32|64 = old behaviour)
mkdir -p gpurun_out
B="python tools/bench_gemm.py --reps 10"
{
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
for d in 96 0 96 0; do echo "== DEBUG=$d"; for s in "res" "k1" "convT" "head 64->64" "256->64" "256->128" "tail2 packed" "dec3 conv1" "expand"; do HIS_GEMM_DEBUG=$d $B --only "$s"; done; done
echo "== stamps"
for s in "256->256 k1 plain" "256->256 k3 64x48 res" "128x96 res"; do HIS_GEMM_DEBUG=16 $B --only "$s"; done
echo "== whole step A/B"
for d in 96 0 96 0; do HIS_GEMM_DEBUG=$d python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('debug=$d', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'])"; done
} > gpurun_out/exp2.log 2>&1
tail -90 gpurun_out/exp2.log
