#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2; do
echo "=== DIRECT=$d"
HIS_GEMM_DIRECT=$d python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline --workload b1 2> gpurun_out/s10_b1_d$d.err | cut -c1-180
grep -E "cin256 cout256 k3|cin72 cout72" gpurun_out/s10_b1_d$d.err | head -4
HIS_GEMM_DIRECT=$d python bench.py --steps 3 --warmup 3 --breakdown --no-cpu-baseline 2> gpurun_out/s10_b0_d$d.err | cut -c1-180
grep -E "cin16 cout16|cin32 cout16" gpurun_out/s10_b0_d$d.err | head -4
done
