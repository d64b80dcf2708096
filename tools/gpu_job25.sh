#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s18_tests.log
tail -5 gpurun_out/s18_tests.log
