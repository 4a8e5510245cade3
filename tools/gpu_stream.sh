#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/manual/stream_check.py "$@" > gpurun_out/stream_check.log 2>&1; echo "exit $?"; tail -40 gpurun_out/stream_check.log
