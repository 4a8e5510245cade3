"""Prints the per-k-block timeline of the w1/w3 GEMM (CTA 0) on the 7B shapes at batch B."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import llama2_ts_b200 as pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hdr = pkg.synth.header("llama2-7b"); hdr[2] = 2          # 2 layers are enough
ctx = pkg.Context(hdr, max_batch=B, max_steps=8)
for t, l, shape in pkg.synth.tensor_plan(hdr):
    a = pkg.synth.gen_tensor_torch(hdr, t, l, 1, "cuda:0").contiguous(); torch.cuda.synchronize(); ctx.upload(t, l, a)
ctx.set_option("graph", 0)
toks = np.arange(2, 2 + B, dtype=np.int32); pos = np.zeros(B, np.int32)
ctx.forward_batch(toks, pos, want_logits=False)
ctx.set_option("gemm_timeline", 1)
ctx.forward_batch(toks, pos + 1, want_logits=False)
tl = ctx.debug_timeline()
t0 = tl[tl > 0].min()
names = ["prod_issue", "split_slotfree", "split_landed", "split_stored", "mma_ready", "mma_issued"]
print("k-block " + " ".join("%14s" % n for n in names))
for n in range(8, 40):
    print("%7d " % n + " ".join("%14d" % (tl[n, i] - t0 if tl[n, i] else -1) for i in range(6)))
d = np.diff(tl[8:60, 5]); print("MMA issue period (cycles):", d[d > 0].mean())
for a, b, nm in ((1, 2, "slot free -> tile landed wait"), (2, 3, "landed -> stored (split work)"), (3, 4, "stored -> MMA sees it"),
                 (4, 5, "MMA issue 8 MMAs + commits")):
    x = tl[8:60, b] - tl[8:60, a]; print("%-34s mean %.0f cycles" % (nm, x[(tl[8:60, a] > 0) & (tl[8:60, b] > 0)].mean()))
