#!/usr/bin/env python
"""Turns gpurun_out/launches.csv (+ prof.ncu-rep) into the tracked summaries under profiles/.
usage: python tools/summarize_profile.py r01"""
import csv
import collections
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
NAMES = {"qkv_attn": "qkv_rope_kvwrite_attention", "<1, 0,": "qkv_rope_kvwrite", "<1, 2,": "w13_swiglu", "<1, 3,": "cls_argmax", "attn_decode": "attention"}


def kname(k, grid_rows=None):
    for pat, n in NAMES.items():
        if pat in k:
            return n
    if "<0, 1," in k:
        return "wo/w2_residual"
    return k[:60]


rows = [r for r in csv.reader(l for l in open(os.path.join(G, "launches.csv")) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
seq = []
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    n = kname(r[ix["Kernel Name"]])
    ns = float(r[ix["Metric Value"]].replace(",", ""))
    if n == "wo/w2_residual":
        n = "w2_residual" if ns > 27000 else "wo_residual"
    per.setdefault(n, []).append(ns)
    seq.append((r[ix["ID"]], n, r[ix["Grid Size"]], r[ix["Block Size"]], ns))
tot = sum(sum(v) for v in per.values())
with open(os.path.join(P, "%s_launches_7b.csv" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): "
            "bench.py --steps 4 --warmup 3, launches 1000..1399 of the gemv/attention kernels\n")
    f.write("id,kernel,grid,block,duration_ns\n")
    for s in seq:
        f.write("%s,%s,\"%s\",\"%s\",%.0f\n" % s)
summ = {k: {"launches": len(v), "avg_us": sum(v) / len(v) / 1e3, "share": sum(v) / tot} for k, v in per.items()}

full = []
rep = os.path.join(G, "prof.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    h, u = rr[0], rr[1]
    ii = {x: i for i, x in enumerate(h)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    for r in rr[2:]:
        d = {"kernel": kname(r[ii["Kernel Name"]]), "kernel_full": r[ii["Kernel Name"]]}
        for w in want:
            if w in ii:
                d[w] = "%s %s" % (r[ii[w]], u[ii[w]])
        full.append(d)
json.dump({"launch_list": summ, "ncu_set_full": full}, open(os.path.join(P, "%s_ncu_summary.json" % tag), "w"), indent=1)
print(json.dumps(summ, indent=1))
