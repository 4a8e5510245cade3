"""Timeline (globaltimer ns) of the GEMV launches of one 7B decode step: for the first and last CTA of
each launch: kernel entry, before/after griddepcontrol.wait, prologue done, warp 0 done, CTA done."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import llama2_ts_b200 as pkg
hdr = pkg.synth.header("llama2-7b"); hdr[2] = 4
ctx = pkg.Context(hdr, max_steps=64)
for t, l, shape in pkg.synth.tensor_plan(hdr):
    a = pkg.synth.gen_tensor_torch(hdr, t, l, 1, "cuda:0").contiguous(); torch.cuda.synchronize(); ctx.upload(t, l, a)
ctx.set_option("graph", 0)
for p in range(40):
    ctx.forward_argmax(5, p)
ctx.set_option("gemv_timeline", 1)
ctx.forward_argmax(5, 40)
tl = ctx.gemv_timeline()
n = int((tl[:, 0, 0] > 0).sum())
t0 = tl[0, 0, 0]
names = ["qkv", "wo", "w13", "w2"]
print("launch kind | CTA0: entry  wait_in wait_out prologue warp0done ctadone | lastCTA: entry wait_out ctadone | gap to next entry")
for i in range(min(n, 17)):
    r0, r1 = tl[i, 0] - t0, tl[i, 1] - t0
    nxt = (tl[i + 1, 0, 0] - t0) if i + 1 < n else -1
    print("%3d %-4s | %7d %7d %7d %7d %7d %7d | %7d %7d %7d | next entry %7d" % (i, names[i % 4] if i < 16 else "cls", *r0, r1[0], r1[2], r1[5], nxt))
d = tl[:16]
print("mean ns: entry->wait_out %.0f, wait_out->prologue %.0f, prologue->ctadone %.0f" % (
    (d[:, 0, 2] - d[:, 0, 0]).mean(), (d[:, 0, 3] - d[:, 0, 2]).mean(), (d[:, 0, 5] - d[:, 0, 3]).mean()))
print("mean ns between CTA0 done of launch i and wait_out of launch i+1: %.0f" % ((d[1:, 0, 2] - d[:-1, 0, 5]).mean()))
print("mean ns spread of ctadone between first and last CTA: %.0f" % (np.abs(d[:, 1, 5] - d[:, 0, 5]).mean()))
