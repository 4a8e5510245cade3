#!/bin/bash
# what the driver does at round end: all GPU tests, smoke, reference arm, our arm
mkdir -p gpurun_out
S=$(date +%s)
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $? ($(( $(date +%s)-S )) s)"; tail -4 gpurun_out/pytest_all.log
S=$(date +%s); timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $? ($(( $(date +%s)-S )) s)"; tail -2 gpurun_out/smoke.log
S=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "reference arm exit $? ($(( $(date +%s)-S )) s)"; tail -c 900 gpurun_out/bench_ref.log; echo
S=$(date +%s); timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $? ($(( $(date +%s)-S )) s)"
python - <<'PY'
import json
for ln in open('gpurun_out/bench.log'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print('7B tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'step frac', round(d['roofline']['step']['frac'],3), 'dom frac', round(d['roofline']['frac'],3), d['clocks'])
        print('cpu', d['cpu_baseline']); print('others', d.get('others')); print('prefill', d.get('prefill')); print('sampling', d.get('sampling')); print('loader', d.get('loader'))
PY
