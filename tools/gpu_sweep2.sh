#!/bin/bash
mkdir -p gpurun_out
for opts in "--opt fuse_prefetch=0" "--opt fuse_prefetch=100" "--opt fuse_prefetch=50" "--opt fuse_prefetch=100 --opt l2_prefetch=0"; do
  timeout 300 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --no-others $opts > gpurun_out/sweep.log 2>&1
  python - "$opts" <<'PY'
import json,sys
for ln in open('gpurun_out/sweep.log'):
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['per_kernel']
        print(sys.argv[1] or 'default', '| tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), '|', ' '.join('%s=%.1f'%(n[:4],v['avg_us']) for n,v in k.items()))
PY
done
