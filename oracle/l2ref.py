"""ctypes wrapper of oracle/l2ref.c -- the CPU restatement of llama2.ts's
transformer() path.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product (llama2.ts_b200/)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libl2ref.so")


def build(force=False):
    src = os.path.join(HERE, "l2ref.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(SO):
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        f32p, i32p, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
        sig = {
            "l2ref_accum": (None, [f32p, f32p, C.c_int]),
            "l2ref_rmsnorm": (None, [f32p, f32p, f32p, C.c_int]),
            "l2ref_softmax": (None, [f32p, C.c_int]),
            "l2ref_matmul": (None, [f32p, f32p, f32p, C.c_int, C.c_int]),
            "l2ref_rope": (None, [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int]),
            "l2ref_set_threads": (None, [C.c_int]),
            "l2ref_max_threads": (C.c_int, []),
            "l2ref_create": (vp, [i32p]),
            "l2ref_weight_floats": (C.c_uint64, [vp]),
            "l2ref_bind_weights": (C.c_int, [vp, f32p]),
            "l2ref_destroy": (None, [vp]),
            "l2ref_logits": (f32p, [vp]),
            "l2ref_x": (f32p, [vp]),
            "l2ref_key_cache": (f32p, [vp]),
            "l2ref_value_cache": (f32p, [vp]),
            "l2ref_vocab": (C.c_int, [vp]),
            "l2ref_forward": (C.c_int, [vp, C.c_int, C.c_int]),
            "l2ref_random_u32": (C.c_uint32, [C.POINTER(C.c_uint64)]),
            "l2ref_random_f32": (C.c_float, [C.POINTER(C.c_uint64)]),
            "l2ref_argmax": (C.c_int, [f32p, C.c_int]),
            "l2ref_sample": (C.c_int, [f32p, C.c_int, C.POINTER(C.c_uint64)]),
            "l2ref_sample_topp": (C.c_int, [f32p, C.c_int, C.c_double, C.POINTER(C.c_uint64)]),
            "l2ref_sample_next": (C.c_int, [f32p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_uint64)]),
            "l2ref_generate": (C.c_int, [vp, C.c_int, i32p, C.c_int, C.c_double, C.c_double,
                                         C.c_uint64, i32p, f32p]),
            "l2ref_tokenizer_load": (vp, [C.c_char_p, C.c_uint64, C.c_int]),
            "l2ref_tokenizer_free": (None, [vp]),
            "l2ref_tokenizer_piece": (C.c_char_p, [vp, C.c_int]),
            "l2ref_tokenizer_score": (C.c_float, [vp, C.c_int]),
            "l2ref_bpe_encode": (C.c_int, [vp, C.c_char_p, i32p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def set_threads(n):
    lib().l2ref_set_threads(int(n))


def max_threads():
    return int(lib().l2ref_max_threads())


# ---- primitives (llama2.ts:165-203) ------------------------------------------
def rmsnorm(x, weight):
    o = np.empty_like(x)
    lib().l2ref_rmsnorm(_f(o), _f(x), _f(weight), x.size)
    return o


def softmax(x):
    y = x.copy()
    lib().l2ref_softmax(_f(y), y.size)
    return y


def matmul(x, w):
    d, n = w.shape
    out = np.empty(d, dtype=np.float32)
    lib().l2ref_matmul(_f(out), _f(x), _f(w), n, d)
    return out


def accum(a, b):
    y = a.copy()
    lib().l2ref_accum(_f(y), _f(b), y.size)
    return y


def rope(q, k, fcr, fci, pos, head_size):
    q2, k2 = q.copy(), k.copy()
    lib().l2ref_rope(_f(q2), _f(k2), _f(fcr), _f(fci), pos, q.size, head_size)
    return q2, k2


class Model:
    """Config + weights + RunState of the reference (llama2.ts:69-163) on the CPU."""

    def __init__(self, hdr, blob):
        """hdr: 7 header ints; blob: float32 array of everything after the header
        (file order, llama2.ts:112-129).  The blob is borrowed, keep it alive."""
        self.hdr = np.asarray(hdr, dtype=np.int32)
        self.blob = np.ascontiguousarray(blob, dtype=np.float32)
        self.h = lib().l2ref_create(_i(self.hdr))
        n = lib().l2ref_weight_floats(self.h)
        assert n == self.blob.size, (n, self.blob.size)
        lib().l2ref_bind_weights(self.h, _f(self.blob))
        self.dim, self.n_layers = int(hdr[0]), int(hdr[2])
        self.vocab = abs(int(hdr[5]))
        self.seq_len = int(hdr[6])

    def __del__(self):
        try:
            if getattr(self, "h", None) and _lib is not None:
                _lib.l2ref_destroy(self.h)
                self.h = None
        except Exception:          # interpreter shutdown
            pass

    def forward(self, token, pos):
        """transformer(token,pos) -> copy of s.logits (llama2.ts:205-303)."""
        rc = lib().l2ref_forward(self.h, int(token), int(pos))
        assert rc == 0, "token/pos out of range"
        return np.ctypeslib.as_array(lib().l2ref_logits(self.h), (self.vocab,)).copy()

    def x(self):
        return np.ctypeslib.as_array(lib().l2ref_x(self.h), (self.dim,)).copy()

    def key_row(self, layer, pos):
        kc = np.ctypeslib.as_array(lib().l2ref_key_cache(self.h), (self.n_layers, self.seq_len, self.dim))
        return kc[layer, pos].copy()

    def value_row(self, layer, pos):
        vc = np.ctypeslib.as_array(lib().l2ref_value_cache(self.h), (self.n_layers, self.seq_len, self.dim))
        return vc[layer, pos].copy()

    def generate(self, steps, prompt=(), temperature=0.0, topp=1.0, seed=1, want_logits=False):
        """The generate loop of llama2.ts:460-508.  Returns (tokens, logits or None)."""
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        if steps <= 0 or steps > self.seq_len:
            steps = self.seq_len
        out = np.zeros(steps, dtype=np.int32)
        lg = np.zeros((steps, self.vocab), dtype=np.float32) if want_logits else None
        n = lib().l2ref_generate(self.h, steps, _i(prompt) if prompt.size else None, prompt.size,
                                 float(temperature), float(topp), int(seed), _i(out),
                                 _f(lg) if want_logits else None)
        assert n >= 0
        return out[:n], (lg[:n] if want_logits else None)


# ---- host pieces -----------------------------------------------------------------
class Rng:
    def __init__(self, seed):
        self.s = C.c_uint64(seed)

    def u32(self):
        return int(lib().l2ref_random_u32(C.byref(self.s)))

    def f32(self):
        return float(lib().l2ref_random_f32(C.byref(self.s)))


def argmax(a):
    return int(lib().l2ref_argmax(_f(a), a.size))


def sample_next(logits, temperature, topp, rng):
    lg = logits.copy()
    return int(lib().l2ref_sample_next(_f(lg), lg.size, temperature, topp, C.byref(rng.s)))


class Tokenizer:
    def __init__(self, path, vocab_size=32000):
        data = open(path, "rb").read()
        self._data = data
        self.h = lib().l2ref_tokenizer_load(data, len(data), vocab_size)

    def encode(self, text):
        out = np.zeros(len(text) + 1, dtype=np.int32)
        n = lib().l2ref_bpe_encode(self.h, text.encode(), _i(out))
        if n < 0:
            raise ValueError("character not found in vocab")
        return out[:n]

    def __del__(self):
        if getattr(self, "h", None):
            lib().l2ref_tokenizer_free(self.h)
            self.h = None
