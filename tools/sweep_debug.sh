#!/bin/bash
# elimination runs of the conv-GEMM roles on one shape: HIS_GEMM_DEBUG bit mask 1 no epilogue stores, 2 no A loads, 4 no B loads, 8 no MMAs
ONLY="${1:-dec4 conv2}"
for dbg in 0 1 2 8 3 10 11; do
  echo "== HIS_GEMM_DEBUG=$dbg"; HIS_GEMM_DEBUG=$dbg python tools/bench_gemm.py --only "$ONLY" --reps 10
done
for nacc in 2 4; do echo "== NACC=$nacc"; HIS_GEMM_NACC=$nacc python tools/bench_gemm.py --only "$ONLY" --reps 10; done
for ast in 2 6 12; do echo "== ASTAGES=$ast"; HIS_GEMM_ASTAGES=$ast python tools/bench_gemm.py --only "$ONLY" --reps 10; done
