P=r2h
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 --breakdown --top 400 > gpurun_out/${P}_bench_b0.json 2> gpurun_out/${P}_bench_b0_breakdown.txt; cut -c1-160 gpurun_out/${P}_bench_b0.json
timeout 150 python bench.py --steps 3 --warmup 3 --precision strict --quick --no-cpu-baseline --breakdown > gpurun_out/${P}_bench_strict.json 2> gpurun_out/${P}_bench_strict_breakdown.txt; cut -c1-160 gpurun_out/${P}_bench_strict.json
timeout 100 python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-graph --no-pipeline > /dev/null 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 624 -c 208 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-graph --no-pipeline > gpurun_out/${P}_ncu_launches.log 2>&1
tail -1 gpurun_out/${P}_ncu_launches.log | cut -c1-120
