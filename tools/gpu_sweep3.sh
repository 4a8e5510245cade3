#!/bin/bash
mkdir -p gpurun_out
for B in 8 16 32; do
for opts in "--opt attn_warp=2048" "--opt attn_warp=1"; do
  timeout 300 python bench.py --workload llama2-7b --batch $B --steps 48 --warmup 4 --no-cpu-baseline --no-others $opts > gpurun_out/sweep.log 2>&1
  python - "B=$B $opts" <<'PY'
import json,sys
for ln in open('gpurun_out/sweep.log'):
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['per_kernel']
        print(sys.argv[1], '| tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), '|', ' '.join('%s=%.1f'%(n[:8],v['avg_us']) for n,v in k.items()))
PY
done
done
