#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s17_b0_2gpu.json 2> gpurun_out/s17_b0_2gpu.err; echo rc=$?; cut -c1-300 gpurun_out/s17_b0_2gpu.json; tail -3 gpurun_out/s17_b0_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s17_ref_2gpu.json 2> gpurun_out/s17_ref_2gpu.err; echo rc=$?; cut -c1-300 gpurun_out/s17_ref_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --workload post --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s17_post_2gpu.json 2> gpurun_out/s17_post_2gpu.err; echo rc=$?; cut -c1-200 gpurun_out/s17_post_2gpu.json
