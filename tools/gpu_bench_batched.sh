#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_$name.log 2>&1; echo "$name exit $?"; tail -c 3000 gpurun_out/bench_$name.log; echo; }
run 110m_b64 --workload stories110M --batch 64 --steps 128 --warmup 8 --no-cpu-baseline
run 7b_b32 --workload llama2-7b --batch 32 --steps 64 --warmup 4 --no-cpu-baseline
run 7b_b256 --workload llama2-7b --batch 256 --steps 64 --warmup 4 --no-cpu-baseline
run 7b_b8 --workload llama2-7b --batch 8 --steps 64 --warmup 4 --no-cpu-baseline
