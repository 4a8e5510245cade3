#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not batch and not prefill and not tc and not tp" > gpurun_out/pytest_soft.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_soft.log
for f in 1 0; do
  timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --no-others --opt soft_sync=$f > gpurun_out/bench_soft$f.log 2>&1
  python - $f <<'PY'
import json,sys
for ln in open('gpurun_out/bench_soft%s.log' % sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['per_kernel']
        print('soft', sys.argv[1], '| tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], '|', ' '.join('%s=%.1f'%(n[:8],v['avg_us']) for n,v in k.items()))
PY
done
timeout 300 python tools/small_sweep.py 2>&1 | grep default
