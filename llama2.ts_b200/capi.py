"""ctypes binding of include/llama2_b200.h -- the same entry points a bun:ffi
``dlopen`` or a Node N-API shim binds.

Nothing here computes: every call goes to libllama2_b200.so, and a missing
library or a missing B200 raises (there is no CPU fallback).
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("L2B_LIBRARY") or os.path.join(HERE, "libllama2_b200.so")   # override: A/B builds
HEADER_PATH = os.path.join(HERE, "..", "include", "llama2_b200.h")

# tensor ids, field order of `interface TransformerWeights` (llama2.ts:95-110)
T_TOKEN_EMBEDDING_TABLE, T_RMS_ATT_WEIGHT, T_WQ, T_WK, T_WV, T_WO, T_RMS_FFN_WEIGHT, \
    T_W1, T_W2, T_W3, T_RMS_FINAL_WEIGHT, T_FREQ_CIS_REAL, T_FREQ_CIS_IMAG, T_WCLS = range(14)
TENSOR_NAMES = ["token_embedding_table", "rms_att_weight", "wq", "wk", "wv", "wo",
                "rms_ffn_weight", "w1", "w2", "w3", "rms_final_weight", "freq_cis_real",
                "freq_cis_imag", "wcls"]
LAYERED = {T_RMS_ATT_WEIGHT, T_WQ, T_WK, T_WV, T_WO, T_RMS_FFN_WEIGHT, T_W1, T_W2, T_W3}

# error codes
OK, EINVAL, EORDER, ECUDA, ESTATE, ENOMEM, ECOMM = 0, -1, -2, -3, -4, -5, -6
# kernel classes / state taps
K_QKV, K_ATTN, K_WO, K_W13, K_W2, K_CLS = 0, 1, 2, 3, 4, 5
K_GEMM_QKV, K_GEMM_WO, K_GEMM_W13, K_GEMM_W2, K_GEMM_CLS, K_BATCH_EPI, K_COUNT = 6, 7, 8, 9, 10, 11, 12
KERNEL_NAMES = ["qkv_rope_kvwrite", "attention", "wo_residual", "w13_swiglu", "w2_residual",
                "cls_argmax", "gemm_qkv", "gemm_wo", "gemm_w13", "gemm_w2", "gemm_cls", "batch_epilogues"]
S_X, S_KEY_ROW, S_VALUE_ROW, S_Q, S_XB, S_HB, S_LOGITS = range(7)


class L2BError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("l2b error %d: %s" % (code, msg))
        self.code = code


def declared_symbols(header_path=HEADER_PATH):
    """Names of every function include/llama2_b200.h declares."""
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(l2b_[a-z0-9_]+)\s*\(", src)))


_i32, _i64, _u64, _f32 = C.c_int32, C.c_int64, C.c_uint64, C.c_float
_p = C.c_void_p
_SIGNATURES = {
    "l2b_create": (C.c_int, [C.POINTER(_i32), _i32, _i32, _i32, C.POINTER(_p)]),
    "l2b_create_tp": (C.c_int, [C.POINTER(_i32), _i32, _i32, _i32, _i32, C.POINTER(_p)]),
    "l2b_create_multi": (C.c_int, [C.POINTER(_i32), _i32, _i32, _i32, _i32, C.POINTER(_p)]),
    "l2b_upload": (C.c_int, [_p, _i32, _i32, _p, _u64]),
    "l2b_load_checkpoint": (C.c_int, [_p, C.c_char_p, C.POINTER(C.c_double)]),
    "l2b_weights_ready": (C.c_int, [_p]),
    "l2b_forward": (C.c_int, [_p, _i32, _i32, _p]),
    "l2b_forward_argmax": (C.c_int, [_p, _i32, _i32, C.POINTER(_i32)]),
    "l2b_forward_batch": (C.c_int, [_p, _i32, _p, _p, _p, _p]),
    "l2b_forward_sample": (C.c_int, [_p, _i32, _i32, C.c_double, C.c_double, _f32, C.POINTER(_i32)]),
    "l2b_sample_logits": (C.c_int, [_p, _p, C.c_double, C.c_double, _f32, C.POINTER(_i32)]),
    "l2b_generate_greedy": (C.c_int, [_p, _i32, _p, _p, _i32, _p, _p]),
    "l2b_prefill": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, C.POINTER(_i32)]),
    "l2b_last_device_ms": (_f32, [_p]),
    "l2b_last_launches": (_i64, [_p]),
    "l2b_profile_step": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "l2b_profile_batch": (C.c_int, [_p, _i32, _p, _p, _p, _p]),
    "l2b_read_state": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _u64]),
    "l2b_reset": (C.c_int, [_p]),
    "l2b_debug_timeline": (C.c_int, [_p, _p, _u64]),
    "l2b_set_option": (C.c_int, [_p, C.c_char_p, _i64]),
    "l2b_tp_export": (_i64, [_p, _p, _u64]),
    "l2b_tp_connect": (C.c_int, [_p, C.c_char_p, _u64, _i32]),
    "l2b_tok_load": (C.c_int, [C.c_char_p, _u64, _i32, C.POINTER(_p)]),
    "l2b_tok_encode": (C.c_int, [_p, C.c_char_p, _p, _i32]),
    "l2b_tok_piece": (C.c_void_p, [_p, _i32]),
    "l2b_tok_piece_len": (_i32, [_p, _i32]),
    "l2b_tok_score": (_f32, [_p, _i32]),
    "l2b_tok_free": (None, [_p]),
    "l2b_last_error": (C.c_char_p, [_p]),
    "l2b_abi_version": (C.c_int, []),
    "l2b_destroy": (None, [_p]),
}


class Library:
    """dlopen of libllama2_b200.so with typed prototypes."""
    _inst = None

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise L2BError(ECUDA, "libllama2_b200.so is not built (%s); run "
                           "`python -c 'import __graft_entry__ as g; g.build()'` -- "
                           "there is no CPU fallback" % path)
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = Library()
        return cls._inst

    def exported(self, name):
        try:
            getattr(self.dll, name)
            return True
        except AttributeError:
            return False


def _ptr(a):
    """Host numpy array, torch tensor (host or device) or raw int address -> void*."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("unsupported buffer %r" % type(a))


class Context:
    """One l2b_ctx: weights + RunState(s) + KV cache on one B200."""

    def __init__(self, hdr, device=0, max_batch=1, max_steps=0, lib=None, tp_rank=0, tp_size=1,
                 n_gpus=0, tp_degree=1):
        """tp_size > 1: this process is rank `tp_rank` of a row-sharded tensor-parallel group
        (l2b_create_tp); call tp_export()/tp_connect() (or dist.connect_tp) before the first step.
        n_gpus >= 1: single-process multi-GPU context (l2b_create_multi) over devices 0..n_gpus-1:
        tp_degree == 1 partitions the max_batch sequences, tp_degree == n_gpus is one
        tensor-parallel group (max_batch 1)."""
        self.lib = lib or Library.get()
        self.hdr = [int(v) for v in hdr]
        assert len(self.hdr) == 7
        self.dim, self.hidden_dim, self.n_layers, self.n_heads = self.hdr[:4]
        self.vocab_size = abs(self.hdr[5])
        self.seq_len = self.hdr[6]
        self.shared_weights = self.hdr[5] > 0
        self.max_batch = max_batch
        self.max_steps = max_steps or self.seq_len
        h = (_i32 * 7)(*self.hdr)
        out = _p()
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.n_gpus = n_gpus
        if n_gpus >= 1:
            rc = self.lib.dll.l2b_create_multi(h, n_gpus, tp_degree, max_batch, max_steps, C.byref(out))
        elif tp_size > 1:
            rc = self.lib.dll.l2b_create_tp(h, device, max_steps, tp_rank, tp_size, C.byref(out))
        else:
            rc = self.lib.dll.l2b_create(h, device, max_batch, max_steps, C.byref(out))
        if rc != 0:
            raise L2BError(rc, self.lib.dll.l2b_last_error(None).decode())
        self._h = out

    # -- plumbing ---------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise L2BError(rc, self.lib.dll.l2b_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dll.l2b_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- weights ----------------------------------------------------------------
    def upload(self, tensor_id, layer, arr, n_floats=None):
        """arr: float32 numpy array / torch tensor (host, or device memory -- the copy is
        cudaMemcpyDefault) holding the tensor exactly as readWeights() slices it."""
        if isinstance(arr, np.ndarray):
            assert arr.dtype == np.float32 and arr.flags.c_contiguous
            n = arr.size
        elif hasattr(arr, "numel"):
            assert arr.is_contiguous() and arr.element_size() == 4
            n = arr.numel()
        else:
            n = n_floats
        self._check(self.lib.dll.l2b_upload(self._h, tensor_id, layer, _ptr(arr), n))

    def load_checkpoint(self, path):
        """Stream a legacy-v0 .bin file into HBM (l2b_load_checkpoint).  Returns seconds."""
        secs = C.c_double(0.0)
        self._check(self.lib.dll.l2b_load_checkpoint(self._h, os.fsencode(path), C.byref(secs)))
        return float(secs.value)

    def weights_ready(self):
        return bool(self.lib.dll.l2b_weights_ready(self._h))

    # -- the hot path -----------------------------------------------------------
    def forward(self, token, pos, logits_out=None):
        """transformer(token,pos,...) -> logits (llama2.ts:468).  Returns float32[vocab]."""
        if logits_out is None:
            logits_out = np.empty(self.vocab_size, dtype=np.float32)
        self._check(self.lib.dll.l2b_forward(self._h, token, pos, _ptr(logits_out)))
        return logits_out

    def forward_argmax(self, token, pos):
        nxt = _i32(0)
        self._check(self.lib.dll.l2b_forward_argmax(self._h, token, pos, C.byref(nxt)))
        return int(nxt.value)

    def forward_sample(self, token, pos, temperature, topp, rand01):
        """One step + temperature/softmax/sampler on the device; rand01 = the host's random_f32()."""
        nxt = _i32(0)
        self._check(self.lib.dll.l2b_forward_sample(self._h, token, pos, float(temperature), float(topp),
                                                    float(rand01), C.byref(nxt)))
        return int(nxt.value)

    def sample_logits(self, logits, temperature, topp, rand01):
        logits = np.ascontiguousarray(logits, dtype=np.float32)
        assert logits.size == self.vocab_size
        nxt = _i32(0)
        self._check(self.lib.dll.l2b_sample_logits(self._h, _ptr(logits), float(temperature), float(topp),
                                                   float(rand01), C.byref(nxt)))
        return int(nxt.value)

    def forward_batch(self, tokens, pos, want_logits=True, want_argmax=True):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        B = tokens.size
        logits = np.empty((B, self.vocab_size), dtype=np.float32) if want_logits else None
        am = np.empty(B, dtype=np.int32) if want_argmax else None
        self._check(self.lib.dll.l2b_forward_batch(self._h, B, _ptr(tokens), _ptr(pos),
                                                   _ptr(logits), _ptr(am)))
        return logits, am

    def prefill(self, tokens, pos0=0, seq=0, want_logits=True):
        """All prompt positions in one pass (l2b_prefill).  Returns (last logits or None, argmax)."""
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        logits = np.empty(self.vocab_size, dtype=np.float32) if want_logits else None
        am = _i32(0)
        self._check(self.lib.dll.l2b_prefill(self._h, seq, tokens.size, _ptr(tokens), pos0, _ptr(logits),
                                             C.byref(am)))
        return logits, int(am.value)

    def generate_greedy(self, tokens, pos, n_steps, forced=None):
        """Device-resident greedy loop; returns int32[n_steps, B] of `next` tokens."""
        tokens = np.ascontiguousarray(np.atleast_1d(tokens), dtype=np.int32)
        pos = np.ascontiguousarray(np.atleast_1d(pos), dtype=np.int32)
        B = tokens.size
        out = np.empty((n_steps, B), dtype=np.int32)
        f = None
        if forced is not None:
            f = np.ascontiguousarray(forced, dtype=np.int32).reshape(n_steps, B)
        self._check(self.lib.dll.l2b_generate_greedy(self._h, B, _ptr(tokens), _ptr(pos), n_steps,
                                                     _ptr(f), _ptr(out)))
        return out

    # -- measurement / debug ------------------------------------------------------
    def last_device_ms(self):
        return float(self.lib.dll.l2b_last_device_ms(self._h))

    def last_launches(self):
        return int(self.lib.dll.l2b_last_launches(self._h))

    def profile_step(self, token, pos):
        ms = np.zeros(K_COUNT, dtype=np.float32)
        n = np.zeros(K_COUNT, dtype=np.int32)
        self._check(self.lib.dll.l2b_profile_step(self._h, token, pos, _ptr(ms), _ptr(n)))
        return ms, n

    def profile_batch(self, tokens, pos):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        ms = np.zeros(K_COUNT, dtype=np.float32)
        n = np.zeros(K_COUNT, dtype=np.int32)
        self._check(self.lib.dll.l2b_profile_batch(self._h, tokens.size, _ptr(tokens), _ptr(pos),
                                                   _ptr(ms), _ptr(n)))
        return ms, n

    def read_state(self, which, seq=0, layer=0, pos=0):
        n = {S_HB: self.hidden_dim, S_LOGITS: self.vocab_size}.get(which, self.dim)
        out = np.empty(n, dtype=np.float32)
        self._check(self.lib.dll.l2b_read_state(self._h, which, seq, layer, pos, _ptr(out), n))
        return out

    def tp_export(self):
        """Opaque handle blob of this rank's exchange block (bytes)."""
        buf = C.create_string_buffer(256)
        n = self.lib.dll.l2b_tp_export(self._h, buf, 256)
        if n < 0:
            self._check(int(n))
        return buf.raw[:n]

    def tp_connect(self, blobs):
        """blobs: list of every rank's tp_export() in rank order."""
        n = len(blobs[0])
        assert all(len(b) == n for b in blobs)
        joined = b"".join(blobs)
        self._check(self.lib.dll.l2b_tp_connect(self._h, joined, n, len(blobs)))

    def gemv_timeline(self):
        out = np.zeros((1024, 2, 6), dtype=np.int64)
        self._check(self.lib.dll.l2b_debug_timeline(self._h, _ptr(out), out.size))
        return out

    def debug_timeline(self):
        out = np.zeros((256, 8), dtype=np.int64)
        self._check(self.lib.dll.l2b_debug_timeline(self._h, _ptr(out), out.size))
        return out

    def reset(self):
        self._check(self.lib.dll.l2b_reset(self._h))

    def set_option(self, key, value):
        self._check(self.lib.dll.l2b_set_option(self._h, key.encode(), int(value)))


class Tokenizer:
    """Native tokenizer.bin + bpe_encode (l2b_tok_*); no device needed."""

    def __init__(self, path_or_bytes, vocab_size=32000, lib=None):
        self.lib = lib or Library.get()
        data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
        h = _p()
        rc = self.lib.dll.l2b_tok_load(bytes(data), len(data), vocab_size, C.byref(h))
        if rc != 0:
            raise L2BError(rc, "malformed tokenizer.bin")
        self._h = h
        self.vocab_size = vocab_size

    def encode(self, text):
        buf = np.zeros(max(1, len(text.encode("utf-8"))), dtype=np.int32)
        n = self.lib.dll.l2b_tok_encode(self._h, text.encode("utf-8"), _ptr(buf), buf.size)
        if n < 0:
            raise ValueError("Error: character not found in vocab")
        return buf[:n].copy()

    def piece(self, i):
        n = self.lib.dll.l2b_tok_piece_len(self._h, int(i))
        return C.string_at(self.lib.dll.l2b_tok_piece(self._h, int(i)), n).decode("utf-8", errors="replace")

    def score(self, i):
        return float(self.lib.dll.l2b_tok_score(self._h, int(i)))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dll.l2b_tok_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
