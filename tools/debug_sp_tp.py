"""Debug aid (2 GPUs): single-process tensor-parallel group vs the one-GPU library on the 7B
architecture, logits step by step, under several launch modes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import llama2_ts_b200 as pkg  # noqa: E402
sys.path.insert(0, ROOT)
from bench import build_weights_on_gpu  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "llama2-7b"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr = pkg.synth.header(arch)
steps = 16
torch.cuda.set_device(0)
one = pkg.Context(hdr, device=0, max_batch=1, max_steps=steps)
build_weights_on_gpu(pkg, one, hdr, 0, "cuda:0")
grp = pkg.Context(hdr, n_gpus=n, tp_degree=n, max_batch=1, max_steps=steps)
build_weights_on_gpu(pkg, grp, hdr, 0, "cuda:0")
one.set_option("fuse_qkv_attn", 0)
for mode in ({}, {"graph": 0}, {"graph": 0, "pdl": 0}):
    for k, v in mode.items():
        grp.set_option(k, v)
    one.reset(); grp.reset()
    tok, bad = 1, None
    for pos in range(steps):
        a = one.forward(tok, pos)
        b = grp.forward(tok, pos)
        if not np.array_equal(a, b) and bad is None:
            bad = (pos, float(np.abs(a - b).max()), int((a != b).sum()))
        tok = int(np.argmax(a))
    one.reset(); grp.reset()
    ta = one.generate_greedy([1], [0], steps)[:, 0]
    tb = grp.generate_greedy([1], [0], steps)[:, 0]
    print(mode, "first logits mismatch:", bad, "| greedy loop equal:", bool(np.array_equal(ta, tb)),
          "| ms/step tp %.3f one %.3f" % (grp.last_device_ms() / steps, one.last_device_ms() / steps), flush=True)
    for k in mode:
        grp.set_option(k, 1)
