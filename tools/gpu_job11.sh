#!/bin/bash
mkdir -p gpurun_out
for n in 2 4; do echo "--- NACC=$n"; HIS_GEMM_NACC=$n timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" 2>&1; done
echo "--- DEBUG=8 (no MMAs)"; HIS_GEMM_DEBUG=8 timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" 2>&1
echo "--- DEBUG=1 (no stores)"; HIS_GEMM_DEBUG=1 timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" 2>&1
echo "--- DEBUG=2 (no A)"; HIS_GEMM_DEBUG=2 timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" 2>&1
echo "--- DEBUG=4 (no B)"; HIS_GEMM_DEBUG=4 timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" 2>&1
echo "--- DEBUG=6 (no A,B)"; HIS_GEMM_DEBUG=6 timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" 2>&1
echo "--- DEBUG=7"; HIS_GEMM_DEBUG=7 timeout 100 python tools/bench_gemm.py --only "128->128 k3 64x48" 2>&1
echo "--- 72: DEBUG sweep"; for d in 0 1 2 4 6 7 8 15; do HIS_GEMM_DEBUG=$d timeout 100 python tools/exp_cin72.py 2>&1 | head -1; done
