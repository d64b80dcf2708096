#!/bin/bash
# round-2 late experiment: residual L2 prefetch A/B + epilogue time stamps (HIS_GEMM_DEBUG=16) on the memory / latency bound shapes
mkdir -p gpurun_out
B="python tools/bench_gemm.py --reps 10"
{
for pf in 0 1 0 1; do echo "== RES_L2PF=$pf"; HIS_GEMM_RES_L2PF=$pf $B --only res; HIS_GEMM_RES_L2PF=$pf $B --only gate; done
echo "== stamps"
for s in "256->256 k1 plain" "258->256" "convT" "gate" "256->256 k3 64x48 res" "head 64->64" "dec4 conv2" "256->64" "128x96 res"; do HIS_GEMM_DEBUG=16 $B --only "$s"; done
echo "== no epilogue stores"
for s in "256->256 k1 plain" "convT" "dec4 conv2" "head 64->64"; do HIS_GEMM_DEBUG=1 $B --only "$s"; done
echo "== whole step A/B"
for pf in 0 1 0 1; do HIS_GEMM_RES_L2PF=$pf python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pf=$pf', d['value'], d['ms_per_step'], d['ms_by_subplan'])"; done
} > gpurun_out/exp1.log 2>&1
tail -60 gpurun_out/exp1.log
