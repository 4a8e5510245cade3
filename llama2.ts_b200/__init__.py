"""llama2.ts_b200 -- B200-native (sm_100a) replacement for ONE path of
wizzard0/llama2.ts: the ``transformer(token, pos, config, state, weights)`` forward
(llama2.ts:205-303, call site llama2.ts:468).

Layout
    csrc/      hand-written CUDA kernels + the C ABI (libllama2_b200.so)
    capi.py    ctypes binding of include/llama2_b200.h (what bun:ffi / N-API binds)
    host.py    host-side mirror of the reference's interface for this path
               (readConfig/readWeights/newRunState/transformer/samplers/main)
    synth.py   named architectures + seeded random-init checkpoints (legacy-v0 .bin)
    build.py   nvcc recipe

The directory name contains a dot, so import it through the root shim:
``import llama2_ts_b200``.  There is no CPU fallback anywhere in this package:
without the built library or without a B200 every compute call raises.
"""
from . import build, capi, dist, host, synth  # noqa: F401
from .capi import L2BError, Library, Context, Tokenizer  # noqa: F401
