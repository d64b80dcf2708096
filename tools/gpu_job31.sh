#!/bin/bash
# source-level profile of the 1x1 expand conv (16->96, SiLU) -- where does the epilogue spend its time?
cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/expand16_96 \
  python tools/bench_gemm.py --only "expand 16->96" --act 2 --reps 1 > gpurun_out/expand_ncu.log 2>&1
tail -3 gpurun_out/expand_ncu.log
ls -la gpurun_out/*.ncu-rep
