#!/bin/bash
# quick A/B: parity subset + default 7B bench (128 steps) + small models
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not batch and not prefill and not tc and not tp and not 7b" > gpurun_out/pytest_ab.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_ab.log
timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --no-others "$@" > gpurun_out/bench_ab.log 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/bench_ab.log'):
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['per_kernel']
        print('tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), '|', ' '.join('%s=%.1f'%(n[:8],v['avg_us']) for n,v in k.items()))
PY
timeout 300 python tools/small_sweep.py 2>&1 | grep default
