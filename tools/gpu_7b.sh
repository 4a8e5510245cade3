#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_7b.py -m gpu -q -x --timeout 900 -s > gpurun_out/pytest_7b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_7b.log
tail -60 gpurun_out/pytest_7b.log
