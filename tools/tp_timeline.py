"""Timeline (globaltimer ns) of rank 0's matvec launches of one tensor-parallel decode step (single-process
group over N GPUs, 7B shapes with 4 layers): kernel entry, griddepcontrol.wait, prologue (= exchange wait +
building the activation vector), streaming, CTA done -- for the first and last CTA of each launch.
    python tools/tp_timeline.py [N]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import llama2_ts_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hdr = pkg.synth.header("llama2-7b"); hdr[2] = 4
torch.cuda.set_device(0)
ctx = pkg.Context(hdr, n_gpus=n, tp_degree=n, max_batch=1, max_steps=64)
for t, l, shape in pkg.synth.tensor_plan(hdr):
    a = pkg.synth.gen_tensor_torch(hdr, t, l, 1, "cuda:0").contiguous(); torch.cuda.synchronize(); ctx.upload(t, l, a)
for graph in (1, 0):
    ctx.set_option("graph", graph)
    for p in range(40 if graph else 8):
        ctx.forward_argmax(5, p if graph else 40 + p)
    ctx.set_option("gemv_timeline", 1)
    ctx.forward_argmax(5, 40 if graph else 48)
    tl = ctx.gemv_timeline()
    ctx.set_option("gemv_timeline", 0)
    k = int((tl[:, 0, 0] > 0).sum())
    if k == 0:
        print("graph=%d: no stamps (kernel nodes were captured before the option was armed)" % graph)
        continue
    t0 = tl[0, 0, 0]
    print("graph=%d  %d launches; per launch, CTA0: entry | wait_out | prologue_done | warp0_done | cta_done ; last CTA: entry, cta_done" % (graph, k))
    for i in range(k):
        r0, r1 = tl[i, 0] - t0, tl[i, 1] - t0
        print("%3d | %7d %7d %7d %7d %7d | %7d %7d" % (i, r0[0], r0[2], r0[3], r0[4], r0[5], r1[0], r1[5]))
    d = tl[1:k - 1]
    print("mean ns: entry->wait_out %.0f, wait_out->prologue %.0f, prologue->cta_done %.0f, cta_done(i)->entry(i+1) %.0f" % (
        (d[:, 0, 2] - d[:, 0, 0]).mean(), (d[:, 0, 3] - d[:, 0, 2]).mean(), (d[:, 0, 5] - d[:, 0, 3]).mean(),
        (tl[2:k, 0, 0] - tl[1:k - 1, 0, 5]).mean()))
