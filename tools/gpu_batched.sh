#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -q -x --timeout 120 -s > gpurun_out/pytest_batched.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_batched.log
tail -40 gpurun_out/pytest_batched.log
