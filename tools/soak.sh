#!/bin/bash
mkdir -p gpurun_out
{
for w in b0 b1 b7 b0_160x120 b0_ln; do timeout 200 python bench.py --steps 25 --warmup 3 --workload $w --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['ms_per_step'], d['steps'], d['clocks']['sm_mhz'])" || echo "$w FAILED rc=$?"; done
timeout 200 python bench.py --steps 12 --warmup 3 --precision strict --quick --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('strict', d['value'], d['ms_per_step'], d['steps'])" || echo "strict FAILED"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} > gpurun_out/soak.log 2>&1
cat gpurun_out/soak.log
