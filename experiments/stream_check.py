"""Streaming kernel (option mega=2): parity against the oracle, then step time per model.

    python tests/manual/stream_check.py [parity] [time] [7b]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
import llama2_ts_b200 as pkg  # noqa: E402
from oracle import l2ref  # noqa: E402

l2ref.build()
what = sys.argv[1:] or ["parity", "time"]


def make(arch, seed, std=0.02, max_steps=0):
    hdr = pkg.synth.header(arch)
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=std)
    ctx = pkg.Context(hdr, device=0, max_batch=1, max_steps=max_steps)
    pkg.synth.upload_blob(ctx, hdr, blob)
    return hdr, blob, ctx


if "parity" in what:
    for arch, seed, std in (("tiny", 1, 0.02), ("tiny-unshared", 2, 0.05), ("small", 3, 0.05), ("stories15M", 11, 0.02),
                            ("wide", 5, 0.02)):
        hdr, blob, ctx = make(arch, seed, std)
        ref = l2ref.Model(hdr, blob)
        S, V = hdr[6], abs(hdr[5])
        n = min(S, 96)
        toks = np.concatenate([[1], pkg.synth.teacher_tokens(n - 1, V, seed)])
        ctx.set_option("mega", 2)
        worst = 0.0
        exact = total = 0
        for pos in range(n):
            got = ctx.forward(int(toks[pos]), pos)
            want = ref.forward(int(toks[pos]), pos)
            d = float(np.abs(got - want).max())
            worst = max(worst, d)
            exact += int((got == want).sum())
            total += V
            if not np.allclose(got, want, rtol=1e-3, atol=1e-4):
                print("  MISMATCH %s pos %d: %g" % (arch, pos, d))
                break
        # greedy loop inside the kernel
        ctx.reset()
        ref2 = l2ref.Model(hdr, blob)
        m = min(S - 1, 64)
        dev = ctx.generate_greedy([1], [0], m)[:, 0]
        tok, want_stream = 1, []
        for pos in range(m):
            lg = ref2.forward(tok, pos)
            tok = l2ref.argmax(lg)
            want_stream.append(tok)
        same = list(dev) == want_stream
        print("%s: max|dlogit| %.3g, bit-identical %.3f%%, greedy stream %s" %
              (arch, worst, 100.0 * exact / total, "identical" if same else "DIFFERENT"))
        if not same:
            print("   dev ", list(dev)[:24])
            print("   want", want_stream[:24])
        ctx.close()

if "time" in what:
    archs = ["stories15M", "stories42M", "stories110M"] + (["llama2-7b"] if "7b" in what else [])
    for arch in archs:
        hdr = pkg.synth.header(arch)
        ctx = pkg.Context(hdr, device=0, max_batch=1, max_steps=0)
        import torch
        for t, l, shape in pkg.synth.tensor_plan(hdr):
            a = pkg.synth.gen_tensor_torch(hdr, t, l, 1, "cuda:0").contiguous()
            torch.cuda.synchronize()
            ctx.upload(t, l, a)
            del a
        torch.cuda.synchronize()
        n = min(hdr[6] - 1, 200)
        res = {}
        for mode in (0, 2):
            ctx.set_option("mega", mode)
            ctx.reset()
            ctx.generate_greedy([1], [0], n)
            ctx.reset()
            t0 = time.time()
            out = ctx.generate_greedy([1], [0], n)[:, 0]
            wall = time.time() - t0
            res[mode] = (ctx.last_device_ms() / n * 1000.0, wall / n * 1e6, list(out))
        print("%s: per-op kernels %.1f us/token (%.0f tok/s) | streaming kernel %.1f us/token (%.0f tok/s) | same stream: %s" %
              (arch, res[0][0], 1e6 / res[0][0], res[2][0], 1e6 / res[2][0], res[0][2] == res[2][2]))
        ctx.close()
