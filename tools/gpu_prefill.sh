#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -s 2>&1 | grep -v "^$" | tail -15
