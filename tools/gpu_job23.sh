#!/bin/bash
mkdir -p gpurun_out
for g in "" "--no-graph"; do echo "=== graph flag: $g"; python bench.py --steps 5 --warmup 3 --no-cpu-baseline $g 2> gpurun_out/s16_b0.err | cut -c1-180; tail -2 gpurun_out/s16_b0.err; done
for g in "" "--no-graph"; do echo "=== graph flag: $g"; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload b7 $g 2> gpurun_out/s16_b7.err | cut -c1-180; tail -2 gpurun_out/s16_b7.err; done
