#!/bin/bash
N=${NG:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --tp --steps 128 --warmup 8 > gpurun_out/bench_7b_tp$N.log 2>&1
echo "tp bench exit $?"; tail -c 2500 gpurun_out/bench_7b_tp$N.log
