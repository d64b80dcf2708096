#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s22_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s22_tests.log
tail -8 gpurun_out/s22_tests.log
