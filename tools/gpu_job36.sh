#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 280 python bench.py --steps 5 --warmup 3 --breakdown --top 12 --no-cpu-baseline > gpurun_out/b0_se.log 2>&1
tail -1 gpurun_out/b0_se.log | cut -c1-260
head -12 gpurun_out/b0_se.log | cut -c1-90
