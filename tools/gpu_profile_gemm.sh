#!/bin/bash
mkdir -p gpurun_out
for PB in 256 32; do
B="python bench.py --workload llama2-7b --batch $PB --steps 3 --warmup 3 --no-cpu-baseline --no-others"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_3xtf32|attn_warp" \
    -s 330 -c 6 -o gpurun_out/prof_gemm$PB -f $B > gpurun_out/ncu_gemm$PB.log 2>&1
echo "ncu $PB exit $?"
done
ls -la gpurun_out/*.ncu-rep
