"""Small end-to-end exercise of every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_run.py [tp]
Runs the default batch-1 path (fused q/k/v+attention, stand-alone kernels), the small-batch
GEMV path, the tcgen05 batched path, prefill, the device sampler -- on `tiny`/`small` shapes so
that a run under the sanitizer takes minutes -- and with `tp` a 2-GPU single-process
tensor-parallel group.  Results are checked against the oracle so a "clean" run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama2_ts_b200 as pkg  # noqa: E402
from oracle import l2ref  # noqa: E402


def check(got, want, what):
    assert np.allclose(got, want, rtol=1e-3, atol=1e-4), (what, float(np.abs(got - want).max()))


def main():
    tp = "tp" in sys.argv[1:]
    light = "light" in sys.argv[1:]          # racecheck is ~100x slower: the tiny shape only
    for arch, steps in ((("tiny", 4),) if light else (("tiny", 6), ("small", 5))):
        hdr = pkg.synth.header(arch)
        _, blob = pkg.synth.checkpoint_blob(hdr, seed=3, std=0.05)
        V = abs(hdr[5])
        toks = np.concatenate([[1], pkg.synth.teacher_tokens(steps - 1, V, 3)])
        ref = l2ref.Model(hdr, blob)
        want = [ref.forward(int(t), p) for p, t in enumerate(toks)]
        if tp:
            with pkg.Context(hdr, n_gpus=2, tp_degree=2, max_batch=1, max_steps=steps) as ctx:
                pkg.synth.upload_blob(ctx, hdr, blob)
                for fuse in (1, 0):
                    ctx.reset()
                    ctx.set_option("fuse_qkv_attn", fuse)
                    for p, t in enumerate(toks):
                        check(ctx.forward(int(t), p), want[p], ("tp", arch, fuse, p))
            print("tp %s ok" % arch)
            continue
        with pkg.Context(hdr, device=0, max_batch=1, max_steps=steps) as ctx:
            pkg.synth.upload_blob(ctx, hdr, blob)
            for opts in ({}, {"fuse_qkv_attn": 0}, {"graph": 0, "pdl": 0}):
                ctx.reset()
                for k, v in opts.items():
                    ctx.set_option(k, v)
                for p, t in enumerate(toks):
                    check(ctx.forward(int(t), p), want[p], (arch, opts, p))
            ctx.reset()
            ctx.generate_greedy([1], [0], steps)
            ctx.reset()
            ctx.forward_sample(1, 0, 0.8, 0.9, 0.4)
            ctx.reset()
            ctx.prefill(toks[:4], 0)
        for B in (2, 5, 33):      # shared-pass GEMV path, tcgen05 path (N = 32 and 64 tiles)
            with pkg.Context(hdr, device=0, max_batch=B, max_steps=steps) as ctx:
                pkg.synth.upload_blob(ctx, hdr, blob)
                for p in range(2):
                    lg, _ = ctx.forward_batch(np.full(B, toks[p], np.int32), np.full(B, p, np.int32))
                    for b in range(B):
                        check(lg[b], want[p], (arch, "B", B, p, b))
        print("%s ok" % arch)


if __name__ == "__main__":
    main()
