#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline --top 400 2> gpurun_out/s15_b0_full.err | cut -c1-180
