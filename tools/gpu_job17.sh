#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post.py -x -q -s > gpurun_out/s12_post_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12_post_tests.log
tail -40 gpurun_out/s12_post_tests.log
