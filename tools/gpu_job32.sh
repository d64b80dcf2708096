#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 280 python bench.py --steps 5 --warmup 3 --breakdown --top 60 --no-cpu-baseline > gpurun_out/b0_new2.log 2>&1
tail -1 gpurun_out/b0_new2.log | cut -c1-300
grep -E "ms_per_step|conv_direct|k1$|depthwise" gpurun_out/b0_new2.log | cut -c1-200 | head -50
HIS_GEMM_DEBUG=0 timeout 120 python tools/bench_gemm.py --only expand --act 2 2>&1 | tail -2
