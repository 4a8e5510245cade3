#!/bin/bash
# ncu launch list + one full capture of the GEMV kernels on the bench workload.
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-others"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv_pairs|attn_decode|qkv_attn" \
    -s 1000 -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?"; tail -3 gpurun_out/ncu_launches.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemv_pairs|attn_decode|qkv_attn" \
    -s 1004 -c 8 -o gpurun_out/prof -f $B > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
