#!/bin/bash
# A/B of the conv kernel's epilogue forms on one box (profiles/r2_exp_epilogue.log): HIS_GEMM_DEBUG=96 selects the general chunk body and
# the wait + barrier at the top of every chunk (the forms before the straight-line body / lean hand-over), 16 prints the epilogue's clock
# stamps when a plan is destroyed.  Every command runs under its own timeout.
mkdir -p gpurun_out
B="timeout 120 python tools/bench_gemm.py --reps 10"
{
for d in 96 0 96 0; do echo "== HIS_GEMM_DEBUG=$d"; for s in "res" "k1" "convT" "head 64->64" "256->64" "expand"; do HIS_GEMM_DEBUG=$d $B --only "$s"; done; done
echo "== stamps"
for s in "256->256 k1 plain" "convT" "gate" "256->256 k3 64x48 res" "head 64->64" "dec4 conv2" "128x96 res"; do HIS_GEMM_DEBUG=16 $B --only "$s"; done
echo "== whole step"
for d in 96 0 96 0; do HIS_GEMM_DEBUG=$d timeout 150 python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('debug=$d', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'])"; done
} > gpurun_out/ab_epilogue.log 2>&1
tail -60 gpurun_out/ab_epilogue.log
