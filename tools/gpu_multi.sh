#!/bin/bash
# N-GPU bench exactly as the driver launches it
N=${NG:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${name}_n$N.log 2>&1; echo "$name exit $?"; tail -c 1500 gpurun_out/bench_${name}_n$N.log; echo; }
run 7b_b1 --steps 128 --warmup 8
run 7b_b256 --batch 256 --steps 64 --warmup 4
