#!/bin/bash
mkdir -p gpurun_out
for a in 2 3 4; do echo "--- ASTAGES=$a"; HIS_GEMM_ASTAGES=$a timeout 200 python tools/bench_gemm.py --only head 2>&1 | tee gpurun_out/s4_sweep_a$a.log; done
timeout 120 python tools/prof_halo.py > gpurun_out/s4_prof_halo_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -o gpurun_out/s4_halo python tools/prof_halo.py > gpurun_out/s4_prof_halo_ncu.log 2>&1
ls -la gpurun_out/s4_halo.ncu-rep
