#!/bin/bash
# one 8-GPU call: partitioned batch-256 at 8 and 4 GPUs, weak batch-1 at 8 (as the driver launches them)
mkdir -p gpurun_out
tr() { N=$1; name=$2; shift 2; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@" > gpurun_out/bench_${name}_n$N.log 2>&1; echo "$name n$N exit $?"; python - gpurun_out/bench_${name}_n$N.log <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); print('   tok/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['scaling'], d['clocks'].get('reasons'))
PY
}
tr 8 7b_b256 --batch 256 --steps 64 --warmup 4 --no-cpu-baseline --no-others
tr 4 7b_b256 --batch 256 --steps 64 --warmup 4 --no-cpu-baseline --no-others
tr 8 7b_b1 --steps 64 --warmup 8 --no-cpu-baseline --no-others
