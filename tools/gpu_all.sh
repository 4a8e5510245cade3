#!/bin/bash
# full GPU regression: all -m gpu tests, then the batched benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -15 gpurun_out/pytest_all.log
