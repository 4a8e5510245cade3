#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_parity.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_parity.log
tail -8 gpurun_out/pytest_parity.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-128} --warmup 8 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
for ln in open('gpurun_out/bench_quick.log'):
    if ln.startswith('{'):
        d=json.loads(ln); print('7B tok/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'step frac', round(d['roofline']['step']['frac'],3)); print({k:round(v['tokens_per_s']) for k,v in d.get('others',{}).items() if 'tokens_per_s' in v})
PY
tail -3 gpurun_out/bench_quick.log | cut -c1-600
