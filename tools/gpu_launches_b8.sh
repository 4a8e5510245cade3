#!/bin/bash
# launch list of one batched step (graph disabled so every kernel is a separate launch)
B=${1:-8}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:bat_|gemm_|attn_|gemv_' --launch-skip 330 -c 330 --csv --log-file gpurun_out/launches_b$B.csv python bench.py --workload llama2-7b --batch $B --steps 2 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/ncu_b$B.log 2>&1; echo exit $?
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_b$B.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(",","")))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print("%-72s n=%4d avg %8.1f us total %9.1f (%.1f%%)" % (k, len(v), sum(v)/len(v)/1000, sum(v)/1000, 100*sum(v)/tot))
PY
