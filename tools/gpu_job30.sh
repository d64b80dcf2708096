#!/bin/bash
# tests + B0 bench with the merged aux conv / head kernel / pool_sum changes, then elimination runs on the 1x1 expand convs
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 280 python bench.py --steps 5 --warmup 3 --breakdown --top 40 --no-cpu-baseline > gpurun_out/b0_new.log 2>&1
tail -1 gpurun_out/b0_new.log | cut -c1-400
for act in 1 2; do
  for dbg in 0 1 2 4 8 15; do
    echo "act=$act debug=$dbg"
    HIS_GEMM_DEBUG=$dbg timeout 120 python tools/bench_gemm.py --only expand --act $act 2>&1 | tail -2
  done
done > gpurun_out/expand_elim.log 2>&1
cat gpurun_out/expand_elim.log
