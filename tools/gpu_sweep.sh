#!/bin/bash
mkdir -p gpurun_out
for opts in "" "--opt attn_prefetch=0" "--opt l2_prefetch=196608" "--opt l2_prefetch=393216" "--opt attn_prefetch=0 --opt l2_prefetch=393216" "--opt threads=256 --opt ctas_per_sm=2"; do
  timeout 300 python bench.py --steps 96 --warmup 8 --no-cpu-baseline --no-others $opts > gpurun_out/sweep.log 2>&1
  python - "$opts" <<'PY'
import json,sys
for ln in open('gpurun_out/sweep.log'):
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['per_kernel']
        print(sys.argv[1] or 'default', '| tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), '|', ' '.join('%s=%.1f'%(n[:4],v['avg_us']) for n,v in k.items()))
PY
done
