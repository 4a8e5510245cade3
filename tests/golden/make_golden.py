#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz by EXECUTING THE REFERENCE'S SOURCE TEXT
(/root/reference/llama2.ts, through oracle/ts_exec.py's mechanical TS->Python
translation with JS number semantics -- the image has no JS engine).  Run it in
the build container (the reference is mounted read-only there):

    python tests/golden/make_golden.py          # golden_v1.npz (tiny shapes, samplers, tokenizer)
    python tests/golden/make_golden.py --v2     # golden_v2.npz (small / wide / stories15M shapes)

The vectors pin oracle/l2ref.c (tests/test_oracle_golden.py) and, through it, the
CUDA path.  Inputs are the seeded synthetic checkpoints of llama2.ts_b200/synth.py;
their SHA-256 is stored so a drifting generator is detected, not silently accepted.
"""
import hashlib
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import llama2_ts_b200 as pkg  # noqa: E402  (synth only: data, no arithmetic)
from oracle import ts_exec  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")


def generate_loop(R, config, state, weights, steps, prompt_tokens, temperature, topp, seed, want_logits):
    """The `while (pos < steps)` loop of llama2.ts:460-508, statement for statement (it lives
    inside main() next to argv/stdout handling, so it is restated here; every function it
    calls is the reference's own translated text)."""
    R["rng_seed"] = seed
    token, pos, out, lg = 1, 0, [], []
    num_prompt_tokens = len(prompt_tokens)
    while pos < steps:
        R["transformer"](token, pos, config, state, weights)             # :468
        if want_logits:
            lg.append(state.logits.a.copy())
        if pos < num_prompt_tokens:                                      # :471
            nxt = int(prompt_tokens[pos])
        else:
            if temperature == 0.0:                                       # :476
                nxt = R["argmax"](state.logits)
            else:
                for q in range(config.vocab_size):                       # :481-483
                    state.logits[q] /= temperature
                R["softmax"](state.logits, 0, config.vocab_size)         # :485
                if topp <= 0 or topp >= 1:                               # :487
                    nxt = R["sample"](state.logits, config.vocab_size)
                else:
                    nxt = R["sample_topp"](state.logits, topp, state.indices)
        pos += 1                                                         # :496
        out.append(int(nxt))
        if nxt == 1:                                                     # :499
            break
        token = nxt
    return np.array(out, dtype=np.int32), (np.array(lg, dtype=np.float32) if want_logits else None)


OUT2 = os.path.join(ROOT, "tests", "golden", "golden_v2.npz")


def forward_vectors(R, g, arch, seed, std, nsteps, keep_kv=True):
    """transformer() (llama2.ts:205-303) of the reference's own text on one seeded synthetic checkpoint:
    logits of every step, final x, the KV rows written."""
    hdr = pkg.synth.header(arch)
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=std)
    key = "%s_s%d" % (arch.replace("-", "_"), seed)
    g[key + "_hdr"] = np.array(hdr, dtype=np.int32)
    g[key + "_std"] = np.array([std])
    g[key + "_sha256"] = np.frombuffer(hashlib.sha256(blob.tobytes()).digest(), dtype=np.uint8)
    config = ts_exec.make_config(hdr)
    weights = ts_exec.make_weights(config, blob)
    state = R["newRunState"](config)
    V = config.vocab_size
    toks = np.concatenate([[1], pkg.synth.teacher_tokens(nsteps - 1, V, seed)]).astype(np.int32)
    lg = []
    for pos in range(nsteps):
        R["transformer"](int(toks[pos]), pos, config, state, weights)
        lg.append(state.logits.a.copy())
        print(key, "pos", pos, flush=True)
    g[key + "_tokens"] = toks
    g[key + "_logits"] = np.array(lg, dtype=np.float32)
    g[key + "_x"] = state.x.a.copy()
    if keep_kv:
        L, S, D = config.n_layers, config.seq_len, config.dim
        g[key + "_key_cache"] = state.key_cache.a.reshape(L, S, D)[:, :nsteps].copy()
        g[key + "_value_cache"] = state.value_cache.a.reshape(L, S, D)[:, :nsteps].copy()
    return key


def main_v2():
    """Round 2: the same pin at the shapes the kernels are specialised for -- `small` (head_size 32, rows of
    one 512-float tile), `wide` (Llama-2-7B's head_size 128, rows of several tiles / K-chunks, hidden size not
    a tile multiple, unshared classifier) and the stories15M shape itself (BASELINE configs[0]: dim 288, 6 heads
    of 48, vocab 32000, shared classifier).  ~4 minutes of CPython."""
    R = ts_exec.load_reference()
    g = {}
    forward_vectors(R, g, "small", 5, 0.05, 8)
    forward_vectors(R, g, "wide", 6, 0.03, 4)
    forward_vectors(R, g, "stories15M", 7, 0.05, 3)
    np.savez_compressed(OUT2, **g)
    print("wrote", OUT2, os.path.getsize(OUT2), "bytes,", len(g), "arrays")


def main():
    R = ts_exec.load_reference()
    F32 = ts_exec.Float32Array
    g = {}
    rng = np.random.default_rng(20231017)

    # ---- primitives (llama2.ts:168-203) -------------------------------------------
    for n in (1, 2, 7, 48, 288):
        x = (rng.standard_normal(n) * 3).astype(np.float32)
        w = (1 + 0.2 * rng.standard_normal(n)).astype(np.float32)
        o = F32(n)
        R["rmsnorm"](o, F32(x.copy()), F32(w.copy()), n)
        g["rmsnorm_%d_x" % n], g["rmsnorm_%d_w" % n], g["rmsnorm_%d_o" % n] = x, w, o.a.copy()
        s = F32(np.concatenate([[9.0], x * 4]).astype(np.float32))      # xPtr = 1
        R["softmax"](s, 1, n)
        g["softmax_%d_x" % n], g["softmax_%d_o" % n] = (x * 4).astype(np.float32), s.a[1:].copy()
        a = F32(x.copy())
        R["accum"](a, F32(w.copy()), n)
        g["accum_%d_o" % n] = a.a.copy()
    for d, n in ((3, 5), (16, 64), (10, 288)):
        W = rng.standard_normal((d, n)).astype(np.float32)
        x = rng.standard_normal(n).astype(np.float32)
        o = F32(d)
        R["matmul"](o, F32(x.copy()), F32(W.ravel().copy()), n, d)
        g["matmul_%dx%d_w" % (d, n)], g["matmul_%dx%d_x" % (d, n)], g["matmul_%dx%d_o" % (d, n)] = W, x, o.a.copy()

    # ---- rng + samplers (llama2.ts:348-394) ------------------------------------------
    for seed in (1, 42, 2**40 + 12345):
        R["rng_seed"] = seed
        g["rng_u32_%d" % seed] = np.array([R["random_u32"]() for _ in range(16)], dtype=np.float64)
        R["rng_seed"] = seed
        g["rng_f32_%d" % seed] = np.array([R["random_f32"]() for _ in range(16)], dtype=np.float32)
    probs_list, choices = [], []
    for trial in range(24):
        V = int(rng.integers(5, 400))
        p = np.exp(rng.standard_normal(V) * rng.uniform(0.5, 4)).astype(np.float32)
        p = (p / p.sum()).astype(np.float32)
        if trial % 4 == 0:
            p[rng.integers(0, V, 3)] = p.max()
        seed = int(rng.integers(1, 2**31))
        row = [V, seed]
        R["rng_seed"] = seed
        row.append(R["sample"](F32(p.copy()), V))
        for topp in (0.9, 0.5, 0.1):
            R["rng_seed"] = seed
            row.append(R["sample_topp"](F32(p.copy()), topp, ts_exec._Array(V)))
        row.append(int(R["argmax"](F32(p.copy()))))
        probs_list.append(np.pad(p, (0, 400 - V)))
        choices.append(row)
    g["sampler_probs"] = np.array(probs_list, dtype=np.float32)
    g["sampler_rows"] = np.array(choices, dtype=np.int64)       # V, seed, sample, topp.9, .5, .1, argmax

    # ---- transformer() on seeded synthetic checkpoints (llama2.ts:205-303) ---------------
    for arch, seed, std, nsteps in (("tiny", 1, 0.02, 12), ("tiny", 21, 0.3, 10), ("tiny-unshared", 2, 0.05, 6)):
        hdr = pkg.synth.header(arch)
        _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=std)
        key = "%s_s%d" % (arch.replace("-", "_"), seed)
        g[key + "_hdr"] = np.array(hdr, dtype=np.int32)
        g[key + "_std"] = np.array([std])
        g[key + "_sha256"] = np.frombuffer(hashlib.sha256(blob.tobytes()).digest(), dtype=np.uint8)
        config = ts_exec.make_config(hdr)
        weights = ts_exec.make_weights(config, blob)
        state = R["newRunState"](config)
        V = config.vocab_size
        toks = np.concatenate([[1], pkg.synth.teacher_tokens(nsteps - 1, V, seed)]).astype(np.int32)
        lg = []
        for pos in range(nsteps):
            R["transformer"](int(toks[pos]), pos, config, state, weights)
            lg.append(state.logits.a.copy())
        g[key + "_tokens"] = toks
        g[key + "_logits"] = np.array(lg, dtype=np.float32)
        g[key + "_x"] = state.x.a.copy()
        L, S, D = config.n_layers, config.seq_len, config.dim
        g[key + "_key_cache"] = state.key_cache.a.reshape(L, S, D)[:, :nsteps].copy()
        g[key + "_value_cache"] = state.value_cache.a.reshape(L, S, D)[:, :nsteps].copy()
        print(key, "forward done", flush=True)
        # generate loops: greedy with a prompt, plain sampling, temperature + top-p
        if arch == "tiny":
            prompt = np.array([17, 300, 45], dtype=np.int32)
            for name, (temp, topp) in (("greedy", (0.0, 1.0)), ("sample", (1.0, 1.0)), ("topp", (0.8, 0.9))):
                state = R["newRunState"](config)
                out, lgs = generate_loop(R, config, state, weights, 14, prompt, temp, topp, 1, name == "greedy")
                g["%s_gen_%s_tokens" % (key, name)] = out
                if lgs is not None:
                    g["%s_gen_%s_logits" % (key, name)] = lgs
            g[key + "_gen_prompt"] = prompt
            print(key, "generate done", flush=True)

    # ---- tokenizer (llama2.ts:441-449, 305-344) on the reference's own tokenizer.bin ---------
    tk = "/root/reference/tokenizer.bin"
    data = open(tk, "rb").read()
    p, vocab, scores = 4, [], []
    for _ in range(32000):
        sc, ln = struct.unpack_from("<fi", data, p)
        p += 8
        vocab.append(data[p:p + ln].decode("utf-8", errors="replace"))   # new TextDecoder().decode
        scores.append(sc)
        p += ln
    prompts = ["Once upon a time", "Hello world", " the quick brown fox", "One day, Lily met a Shoggoth",
               "a", "I believe the meaning of life is"]
    ids = []
    for s in prompts:
        toks = ts_exec.Int32Array(len(s) + 1)
        n = R["bpe_encode"](s, vocab, scores, 32000, toks)
        ids.append(np.pad(toks.a[:n], (0, 64 - n), constant_values=-1))
    g["bpe_prompts"] = np.array(prompts)
    g["bpe_ids"] = np.array(ids, dtype=np.int32)
    g["tokenizer_sha256"] = np.frombuffer(hashlib.sha256(data).digest(), dtype=np.uint8)

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    if "--v2" in sys.argv:
        main_v2()
    else:
        main()
