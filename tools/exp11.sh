#!/bin/bash
mkdir -p gpurun_out
B="timeout 120 python tools/bench_gemm.py --reps 10"
{
for d in 0 1 0 1; do echo "== K1_SPLITN=$d"; for s in "k1" "convT"; do HIS_GEMM_K1_SPLITN=$d $B --only "$s"; done; done
for d in 0 1 0 1; do HIS_GEMM_K1_SPLITN=$d timeout 150 python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('splitn=$d', d['value'], d['ms_per_step'], d['ms_by_subplan'], d['roofline']['frac'])"; done
} > gpurun_out/exp11.log 2>&1
tail -40 gpurun_out/exp11.log
