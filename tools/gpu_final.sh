#!/bin/bash
# Measurement pass of round 2: tests, the bench line (all sub-records) + reference arm, per-workload breakdowns, ncu launch list +
# DRAM traffic of one B0 step, ncu --set full of the kernels under study.  Everything lands in gpurun_out/${P}_* (P = prefix).
P=${1:-r2g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_tests.log; tail -3 gpurun_out/${P}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --breakdown --top 400 > gpurun_out/${P}_bench_b0.json 2> gpurun_out/${P}_bench_b0_breakdown.txt; cut -c1-160 gpurun_out/${P}_bench_b0.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> /dev/null; cut -c1-200 gpurun_out/${P}_bench_ref.json
for w in b1 b7 b0_160x120 b0_ln; do timeout 300 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline --breakdown > gpurun_out/${P}_bench_$w.json 2> gpurun_out/${P}_bench_${w}_breakdown.txt; cut -c1-160 gpurun_out/${P}_bench_$w.json; done
timeout 300 python bench.py --steps 3 --warmup 3 --precision strict --quick --no-cpu-baseline --breakdown > gpurun_out/${P}_bench_strict.json 2> gpurun_out/${P}_bench_strict_breakdown.txt; cut -c1-160 gpurun_out/${P}_bench_strict.json
timeout 300 python bench.py --workload post --steps 5 --warmup 3 > gpurun_out/${P}_bench_post.json 2> /dev/null; cut -c1-160 gpurun_out/${P}_bench_post.json
# one step under ncu: launch durations + DRAM bytes (graph replay off so that every kernel is a separate launch record)
timeout 300 python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-graph --no-pipeline > /dev/null 2>&1 && \
L=$(python -c "import json; print(json.load(open('gpurun_out/${P}_bench_b0.json'))['launches_per_step'])"); echo "launches per step: $L"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-graph --no-pipeline > gpurun_out/${P}_ncu_launches.log 2>&1
tail -1 gpurun_out/${P}_ncu_launches.log | cut -c1-200
timeout 300 python bench.py --workload post --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/${P}_post_launches.csv python bench.py --workload post --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${P}_ncu_post.log 2>&1
timeout 300 python tools/prof_set.py > gpurun_out/${P}_prof_set_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"depthwise|conv_gemm|mask_cleanup|roi_align|dilate_logits|paste_kernel" -c 24 -o gpurun_out/${P}_prof_set python tools/prof_set.py > gpurun_out/${P}_prof_set_ncu.log 2>&1
tail -2 gpurun_out/${P}_prof_set_ncu.log
# the report itself can exceed what gpurun copies back (64 MiB for the whole directory): keep its raw page as CSV, drop the file when large
ncu -i gpurun_out/${P}_prof_set.ncu-rep --page raw --csv > gpurun_out/${P}_prof_set_raw.csv 2> /dev/null
if [ $(stat -c %s gpurun_out/${P}_prof_set.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/${P}_prof_set.ncu-rep; fi
ls -la gpurun_out/${P}_*
