#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown --no-pipeline $W 2> gpurun_out/tmp.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   value %.0f  ms %.2f' % (d['value'], d['ms_per_step']))"; head -1 gpurun_out/tmp.err; }
for rep in 1 2; do
for mx in 64 96 128; do echo "== b0 HALO_MAXN=$mx"; W="" HIS_GEMM_HALO_MAXN=$mx run; done
done
for mx in 64 96 128; do echo "== b1 HALO_MAXN=$mx"; W="--workload b1" HIS_GEMM_HALO_MAXN=$mx run; done
for mx in 64 96 128; do echo "== b7 HALO_MAXN=$mx"; W="--workload b7" HIS_GEMM_HALO_MAXN=$mx run; done
