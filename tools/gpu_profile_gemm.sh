#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload llama2-7b --batch ${PB:-32} --steps 3 --warmup 3 --no-cpu-baseline --no-others"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_3xtf32" \
    -s 330 -c 5 -o gpurun_out/prof_gemm${PB:-32} -f $B > gpurun_out/ncu_gemm.log 2>&1
echo "full exit $?"; tail -3 gpurun_out/ncu_gemm.log
