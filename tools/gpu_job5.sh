#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_tests2.log
tail -30 gpurun_out/s3_tests2.log
