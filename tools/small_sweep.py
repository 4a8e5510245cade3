"""Option sweep for the small models (batch-1 greedy, device-resident loop)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
import llama2_ts_b200 as pkg
import torch
for arch in ("stories15M", "stories42M", "stories110M"):
    hdr = pkg.synth.header(arch)
    ctx = pkg.Context(hdr, device=0, max_batch=1, max_steps=0)
    for t, l, shape in pkg.synth.tensor_plan(hdr):
        a = pkg.synth.gen_tensor_torch(hdr, t, l, 1, "cuda:0").contiguous()
        torch.cuda.synchronize(); ctx.upload(t, l, a); del a
    torch.cuda.synchronize()
    n = min(hdr[6] - 1, 200)
    base = {"threads": 512, "ctas_per_sm": 1, "l2_prefetch": 262144, "evict_first": -1, "attn_cluster": 0, "pdl": 1}
    base.update({"fuse_cluster": 0, "fuse_qkv_attn": 1})
    for opts in ({}, {"fuse_cluster": 4}, {"fuse_cluster": 2}, {"fuse_cluster": 8}, {"fuse_qkv_attn": 0}, {"threads": 256}, {"threads": 256, "ctas_per_sm": 2}, {"l2_prefetch": 0}, {"l2_prefetch": 65536},
                 {"attn_cluster": 1}, {"attn_cluster": 2}, {"attn_cluster": 4}, {"pdl": 0}, {"threads": 256, "l2_prefetch": 0}):
        cfg = dict(base); cfg.update(opts)
        for k, v in cfg.items():
            ctx.set_option(k, v)
        ctx.reset(); ctx.generate_greedy([1], [0], n)
        best = 1e9
        for rep in range(3):
            ctx.reset(); ctx.generate_greedy([1], [0], n)
            best = min(best, ctx.last_device_ms() / n * 1000.0)
        print("%-12s %-44s %.1f us/token (%.0f tok/s)" % (arch, opts or "default", best, 1e6 / best), flush=True)
    ctx.close()
