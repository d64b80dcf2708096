#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown --no-pipeline 2> gpurun_out/tmp.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   value %.0f  ms %.2f' % (d['value'], d['ms_per_step']))"; head -1 gpurun_out/tmp.err; }
for rep in 1 2; do
for mx in 64 128 256; do echo "== HALO_MAXN=$mx"; HIS_GEMM_HALO_MAXN=$mx run; done
echo "== TAPS=1"; HIS_GEMM_TAPS=1 run
echo "== NACC=2"; HIS_GEMM_NACC=2 run
echo "== BRES=0"; HIS_GEMM_BRES=0 run
echo "== FUSE_UP=0"; HIS_GEMM_FUSE_UP=0 run
done
