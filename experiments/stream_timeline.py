"""Per-phase globaltimer stamps of the streaming kernel's last step (CTA 0 and the last CTA)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
import llama2_ts_b200 as pkg
import torch
arch = sys.argv[1] if len(sys.argv) > 1 else "stories15M"
hdr = pkg.synth.header(arch)
ctx = pkg.Context(hdr, device=0, max_batch=1, max_steps=0)
for t, l, shape in pkg.synth.tensor_plan(hdr):
    a = pkg.synth.gen_tensor_torch(hdr, t, l, 1, "cuda:0").contiguous()
    torch.cuda.synchronize(); ctx.upload(t, l, a); del a
torch.cuda.synchronize()
ctx.set_option("mega", 2)
ctx.generate_greedy([1], [0], 20)
ctx.reset()
ctx.set_option("gemv_timeline", 1)
ctx.generate_greedy([1], [0], 20)
print("us/token", ctx.last_device_ms() / 20 * 1000)
tl = ctx.gemv_timeline()[:, :, 0]
names = ["start", "rms", "qkv", "attn", "g_att", "wo", "rms2", "w13", "g_hb", "w2"]
L = hdr[2]
for which in (0, 1):
    print("cta", "0" if which == 0 else "last")
    t0 = tl[0, which]
    prev = t0
    for l in range(min(L, 3)):
        row = []
        for k in range(1, 10):
            t = tl[l * 9 + k, which]
            if t:
                row.append("%s %.1f" % (names[k], (t - prev) / 1000.0)); prev = t
            else:
                row.append("%s -" % names[k])
        print("  layer", l, " | ".join(row))
    tend = tl[L * 9, which]
    print("  all layers: %.1f us" % ((tend - t0) / 1000.0))
