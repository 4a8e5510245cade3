#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x --timeout 400 -s > gpurun_out/pytest_tp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tp.log
tail -30 gpurun_out/pytest_tp.log
