"""Host-side mirror of llama2.ts's interface for the transformer() path.

The image has no JavaScript runtime (node/bun/deno absent), so the host side
above the C ABI is written in Python with the SAME names, argument meaning and
error behaviour as the reference's TypeScript (file:line cited per function).
What the reference keeps on the host stays on the host here too (CLI flags,
tokenizer, xorshift RNG, samplers, generate loop); ``transformer()`` and the
data it owns (weights, RunState, KV cache) go through libllama2_b200.so.
INTEGRATION.md shows the equivalent bun:ffi / N-API patch of llama2.ts itself.

No arithmetic of the hot path happens here and nothing falls back to the CPU.
"""
import math
import struct
import sys
import time
from types import SimpleNamespace

import numpy as np

from . import capi

_MASK64 = (1 << 64) - 1


# ----------------------------------------------------------------------------
# config + weights + state (llama2.ts:69-163)

def readConfig(buf):
    """llama2.ts:80-93.  buf: the first 28 bytes of the checkpoint."""
    dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len = struct.unpack("<7i", buf[:28])
    c = SimpleNamespace()
    c.dim, c.hidden_dim, c.n_layers, c.n_heads, c.n_kv_heads = dim, hidden_dim, n_layers, n_heads, n_kv_heads
    c.vocab_size = abs(vocab_size)
    c.seq_len = seq_len
    c.shared_weights = vocab_size > 0
    c.head_size = dim // n_heads
    c.header = [dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len]
    return c


class TransformerWeights:
    """llama2.ts:95-110, but the tensors live in HBM: this object only owns the device
    context they were uploaded into."""

    def __init__(self, ctx):
        self.ctx = ctx


def readWeights(config, f, shared_weights, device=0, max_steps=0, max_batch=1):
    """llama2.ts:112-129: sequential read in file order; every slice is handed to the
    library right after it is read (l2b_upload copies it into HBM) and then dropped.
    f: binary file object positioned after the 28-byte header."""
    ctx = capi.Context(config.header, device=device, max_batch=max_batch, max_steps=max_steps)
    D, F, L, V, S = config.dim, config.hidden_dim, config.n_layers, config.vocab_size, config.seq_len
    hs2 = config.head_size // 2

    def get(*dims):
        n = int(np.prod(dims))
        a = np.fromfile(f, dtype="<f4", count=n)
        if a.size != n:
            raise EOFError("checkpoint truncated")
        return a

    ctx.upload(capi.T_TOKEN_EMBEDDING_TABLE, 0, get(V, D))
    for t, dims in ((capi.T_RMS_ATT_WEIGHT, (D,)), (capi.T_WQ, (D, D)), (capi.T_WK, (D, D)),
                    (capi.T_WV, (D, D)), (capi.T_WO, (D, D)), (capi.T_RMS_FFN_WEIGHT, (D,)),
                    (capi.T_W1, (F, D)), (capi.T_W2, (D, F)), (capi.T_W3, (F, D))):
        for l in range(L):
            ctx.upload(t, l, get(*dims))
    ctx.upload(capi.T_RMS_FINAL_WEIGHT, 0, get(D))
    ctx.upload(capi.T_FREQ_CIS_REAL, 0, get(S, hs2))
    ctx.upload(capi.T_FREQ_CIS_IMAG, 0, get(S, hs2))
    if not shared_weights:
        ctx.upload(capi.T_WCLS, 0, get(V, D))       # llama2.ts:127
    return TransformerWeights(ctx)


def newRunState(config):
    """llama2.ts:147-163.  Only `logits` (what the host reads, :478-492) and the sampler
    scratch exist on the host; x/xb/q/k/v/att/hb and the KV cache are device memory."""
    s = SimpleNamespace()
    s.logits = np.zeros(config.vocab_size, dtype=np.float32)
    s.indices = None
    return s


def transformer(token, pos, p, s, w):
    """llama2.ts:205-303 -> one l2b_forward call; s.logits receives the vocab_size logits."""
    w.ctx.forward(int(token), int(pos), s.logits)


# ----------------------------------------------------------------------------
# tokenizer (llama2.ts:441-449, 305-344) -- host-side, unchanged semantics

def read_tokenizer(path, vocab_size):
    data = open(path, "rb").read()
    p = 4  # ignored max_token_length
    vocab, scores = [], []
    for _ in range(vocab_size):
        score, ln = struct.unpack_from("<fi", data, p)
        p += 8
        vocab.append(data[p:p + ln].decode("utf-8", errors="replace"))
        scores.append(score)
        p += ln
    return vocab, scores


def bpe_encode(text, vocab, vocab_scores, vocab_size, tokens):
    """llama2.ts:305-344 (indexOf = first match wins)."""
    first = {}
    for i, v in enumerate(vocab):
        first.setdefault(v, i)
    n_tokens = 0
    for ch in text:
        idx = first.get(ch, -1)
        if idx == -1:
            raise ValueError("Error: character not found in vocab: " + ch)
        tokens[n_tokens] = idx
        n_tokens += 1
    while True:
        best_score, best_id, best_idx = -1e10, -1, -1
        for i in range(n_tokens - 1):
            idx = first.get(vocab[tokens[i]] + vocab[tokens[i + 1]], -1)
            if idx != -1 and vocab_scores[idx] > best_score:
                best_score, best_id, best_idx = vocab_scores[idx], idx, i
        if best_idx == -1:
            break
        tokens[best_idx] = best_id
        for i in range(best_idx + 1, n_tokens - 1):
            tokens[i] = tokens[i + 1]
        n_tokens -= 1
    return n_tokens


# ----------------------------------------------------------------------------
# rng + samplers (llama2.ts:346-394) -- host-side so the random stream is unchanged

class Rng:
    def __init__(self, seed):
        self.seed = int(seed) & _MASK64

    def random_u32(self):
        """llama2.ts:349-354 xorshift64*."""
        s = self.seed
        s ^= s >> 12
        s ^= (s << 25) & _MASK64
        s ^= s >> 27
        self.seed = s
        return ((s * 0x2545F4914F6CDD1D) >> 32) & 0xFFFFFFFF

    def random_f32(self):
        """llama2.ts:356-360."""
        return float(np.float32((self.random_u32() / 256) / 16777216.0))


def softmax(x, xPtr, size):
    """llama2.ts:181-194 on a float32 array, in place (host copy used at :485)."""
    v = x[xPtr:xPtr + size]
    m = float(v.max())
    e = np.exp(v.astype(np.float64) - m).astype(np.float32)
    total = float(np.cumsum(e.astype(np.float64))[-1])          # sequential f64 sum
    v[:] = (e.astype(np.float64) / total).astype(np.float32)


def argmax(arr):
    """llama2.ts:364-366: first maximum wins, NaN never wins (NaN at 0 keeps 0)."""
    if arr[0] != arr[0]:
        return 0
    a = np.where(np.isnan(arr), -np.inf, arr)
    return int(np.argmax(a))


def sample(logits, vocabSize, rng):
    """llama2.ts:368-376."""
    cum = np.cumsum(logits[:vocabSize].astype(np.float64))
    r = rng.random_f32() * float(cum[-1])
    i = int(np.searchsorted(cum, r, side="right"))
    return i if i < vocabSize else 0


def sample_topp(logits, topp, rng):
    """llama2.ts:378-394, including the exclusive `i < lastIdx` walk and the fallback 0."""
    order = np.argsort(-logits, kind="stable")
    cum = np.cumsum(logits[order].astype(np.float64))
    over = np.nonzero(cum > topp)[0]
    if over.size:
        lastIdx = int(over[0])
        cumProb = float(cum[lastIdx])
    else:
        lastIdx, cumProb = 0, float(cum[-1])
    r = rng.random_f32() * cumProb
    hit = np.nonzero(r < cum[:lastIdx])[0]
    return int(order[hit[0]]) if hit.size else 0


# ----------------------------------------------------------------------------
# generate loop (llama2.ts:460-511)

def generate(config, weights, state, steps, prompt_tokens, temperature, topp, rng, on_token=None,
             device_greedy=False, prefill=False, device_sampler=False):
    """The `while (pos < steps)` loop.  Returns (tokens, tok_per_s).
    device_greedy=True keeps the whole -t 0 loop on the device (l2b_generate_greedy):
    same tokens, no per-token round trip."""
    if steps <= 0 or steps > config.seq_len:
        steps = config.seq_len                                   # llama2.ts:439
    out = []
    n_prompt = len(prompt_tokens)
    if device_greedy and temperature == 0.0:
        forced = np.full(steps, -1, dtype=np.int32)
        forced[:min(n_prompt, steps)] = prompt_tokens[:steps]
        t0 = time.time()
        nxt = weights.ctx.generate_greedy([1], [0], steps, forced)[:, 0]
        dt = time.time() - t0
        for t in nxt:
            out.append(int(t))
            if t == 1:
                break
        prev = 1
        for t in out:
            if on_token and t != 1:
                on_token(t, prev)
            prev = t
        return out, (len(out) / dt if dt > 0 else float("inf"))
    token, pos, start = 1, 0, 0.0
    if prefill and n_prompt > 1 and n_prompt < steps:
        # positions 0..n_prompt-1 only force the next prompt token (llama2.ts:471-473): run them
        # as ONE batched pass (l2b_prefill) instead of n_prompt transformer() calls
        fed = np.concatenate([[1], np.asarray(prompt_tokens[:n_prompt - 1], dtype=np.int32)])
        weights.ctx.prefill(fed, 0, want_logits=False)
        for i in range(n_prompt):
            nxt = int(prompt_tokens[i])
            out.append(nxt)
            if nxt == 1:
                return out, float("inf")
            if on_token:
                on_token(nxt, token)
            token = nxt
        pos = n_prompt
        start = time.time()
    while pos < steps:
        if device_sampler and pos >= n_prompt and temperature != 0.0:
            # llama2.ts:468 + :481-494 in one call; the random number is still drawn here
            nxt = weights.ctx.forward_sample(token, pos, temperature, topp, rng.random_f32())
        else:
            transformer(token, pos, config, state, weights)      # llama2.ts:468
        if device_sampler and pos >= n_prompt and temperature != 0.0:
            pass
        elif pos < n_prompt:
            nxt = int(prompt_tokens[pos])
        elif temperature == 0.0:
            nxt = argmax(state.logits)
        else:
            state.logits[:] = (state.logits.astype(np.float64) / temperature).astype(np.float32)
            softmax(state.logits, 0, config.vocab_size)
            if topp <= 0 or topp >= 1:
                nxt = sample(state.logits, config.vocab_size, rng)
            else:
                nxt = sample_topp(state.logits, topp, rng)
        pos += 1
        out.append(nxt)
        if nxt == 1:
            break
        if on_token:
            on_token(nxt, token)
        token = nxt
        if start == 0.0:
            start = time.time()
    dt = time.time() - start
    return out, ((pos - 1) / dt if dt > 0 else float("inf"))


def error_usage():
    """llama2.ts:514-524."""
    e = sys.stderr
    print("Usage: ... llama2.ts <checkpoint> [options]", file=e)
    print('Example: llama2.ts model.bin -n 256 -i "Once upon a time"', file=e)
    print("Options:", file=e)
    print("  -t <float>  temperature, default 1.0", file=e)
    print("  -p <float>  p value in top-p (nucleus) sampling. default 0.9, 0 = off", file=e)
    print("  -s <int>    random seed, default time(NULL)", file=e)
    print("  -n <int>    number of steps to run for, default 256. 0 = max_seq_len", file=e)
    print("  -i <string> input prompt", file=e)
    sys.exit(1)


def main(argv, tokenizer_path="tokenizer.bin"):
    """llama2.ts:399-512: same flags, same defaults, same output format."""
    if len(argv) < 1:
        error_usage()
    checkpoint, args = argv[0], argv[1:]
    temperature, topp, seed, steps, prompt = 1.0, 1.0, 0, 256, None
    for i in range(0, len(args), 2):
        if i + 1 >= len(args) or args[i][0] != "-" or len(args[i]) != 2:
            error_usage()
        a, val = args[i][1], args[i + 1]
        if a == "t": temperature = float(val)
        elif a == "p": topp = float(val)
        elif a == "s": seed = int(val)
        elif a == "n": steps = int(val)
        elif a == "i": prompt = val
        else: error_usage()
    if seed == 0:
        seed = int(time.time() * 1000)
    with open(checkpoint, "rb") as f:
        config = readConfig(f.read(28))
        if steps <= 0 or steps > config.seq_len:
            steps = config.seq_len
        weights = readWeights(config, f, config.shared_weights, max_steps=steps)
    vocab, vocab_scores = read_tokenizer(tokenizer_path, config.vocab_size)
    state = newRunState(config)
    prompt_tokens = np.zeros(config.seq_len, dtype=np.int32)
    n_prompt = 0
    if prompt is not None:
        n_prompt = bpe_encode(prompt, vocab, vocab_scores, config.vocab_size, prompt_tokens)

    def emit(nxt, token):
        s = vocab[nxt]
        if token == 1 and s[:1] == " ":
            s = s[1:]                                            # llama2.ts:502
        sys.stdout.write(s)
        sys.stdout.flush()

    _, tps = generate(config, weights, state, steps, prompt_tokens[:n_prompt], temperature, topp,
                      Rng(seed), on_token=emit)
    print("\n\nachieved tok/s: %f\n" % tps)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
