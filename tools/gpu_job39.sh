#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 280 python bench.py --steps 5 --warmup 3 --breakdown --top 5 --no-cpu-baseline > gpurun_out/b0_up.log 2>&1
tail -1 gpurun_out/b0_up.log | cut -c1-200
head -8 gpurun_out/b0_up.log | cut -c1-90
