"""oracle/ts_exec.py -- run the reference's OWN source text without a JS engine.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py to pin oracle/l2ref.c).

The build image has no node/bun/deno, so /root/reference/llama2.ts cannot be
executed natively.  This module reads that file, cuts out the functions on the
transformer() path (by name), translates each one statement by statement into
Python, and executes the result with JavaScript number semantics:

  * every JS `number` is a Python float/int (IEEE double; ints stay exact),
  * `Float32Array` element stores round to float32 (numpy cast, RNE) and loads
    widen back to double -- exactly what a typed array does,
  * `/` is true division, `Math.sqrt` is correctly rounded, `Math.exp` is libm's
    exp (V8 uses an fdlibm port; both are < 1 ulp and the result is rounded to
    float32 immediately at llama2.ts:187 and :285),
  * BigInt arithmetic (the xorshift RNG) is Python int arithmetic,
  * Array.prototype.sort is stable (ES2019), like list.sort.

The translation is mechanical (token/regex rewriting of the reference's own
lines; no arithmetic is re-authored here), so the statements that execute are
the reference's statements.  It handles exactly the subset of TypeScript those
functions use and raises on anything else.
"""
import functools
import math
import re
from types import SimpleNamespace

import numpy as np

REFERENCE = "/root/reference/llama2.ts"
FUNCTIONS = ["newRunState", "accum", "rmsnorm", "softmax", "matmul", "transformer", "bpe_encode",
             "random_u32", "random_f32", "argmax", "sample", "sample_topp"]
GLOBALS_ASSIGNED = {"rng_seed"}


# ---- JS runtime shims -----------------------------------------------------------
class Float32Array:
    def __init__(self, a, _view=None):
        if _view is not None:
            self.a = _view
        elif isinstance(a, (int, float)):
            self.a = np.zeros(int(a), dtype=np.float32)
        else:
            self.a = np.asarray(a, dtype=np.float32)

    @staticmethod
    def _ix(i):
        j = int(i)
        if j != i:
            raise IndexError("non-integral typed-array index %r" % (i,))
        return j

    def __getitem__(self, i):
        return float(self.a[self._ix(i)])          # widen f32 -> double

    def __setitem__(self, i, v):
        self.a[self._ix(i)] = np.float32(v)        # round double -> f32 (RNE)

    def __len__(self):
        return self.a.size

    def subarray(self, begin, end=None):
        b = self._ix(begin)
        return Float32Array(None, _view=self.a[b:] if end is None else self.a[b:self._ix(end)])

    def set(self, src, offset=0):
        o = self._ix(offset)
        self.a[o:o + len(src)] = src.a

    def fill(self, v, begin=0, end=None):
        self.a[self._ix(begin):None if end is None else self._ix(end)] = np.float32(v)

    def reduce(self, fn, init):
        acc = init
        n = fn.__code__.co_argcount
        for idx in range(self.a.size):
            val = float(self.a[idx])
            acc = fn(acc, val, idx, self) if n == 4 else fn(acc, val)
        return acc


class Int32Array(Float32Array):
    def __init__(self, a):
        self.a = np.zeros(int(a), dtype=np.int32) if isinstance(a, (int, float)) else np.asarray(a, dtype=np.int32)

    def __getitem__(self, i):
        return int(self.a[self._ix(i)])

    def __setitem__(self, i, v):
        self.a[self._ix(i)] = int(v)


class JSArray(list):
    def sort(self, cmp=None):
        list.sort(self, key=functools.cmp_to_key(cmp))


def _Array(n):
    return JSArray([None] * int(n))


def _indexOf(seq, item):
    try:
        return seq.index(item)
    except ValueError:
        return -1


def _obj(**kw):
    return SimpleNamespace(**kw)


class _Math:
    sqrt = staticmethod(math.sqrt)
    exp = staticmethod(math.exp)
    abs = staticmethod(abs)


# ---- source extraction ------------------------------------------------------------
def _strip_comments(src):
    return re.sub(r"//[^\n]*", "", src)


def _match(src, i, open_ch, close_ch):
    depth = 0
    while True:
        c = src[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1


def extract_function(src, name):
    """(param_names, body_text) of `function name(...)...{ body }`."""
    m = re.search(r"\bfunction\s+%s\s*\(" % re.escape(name), src)
    if not m:
        raise KeyError(name)
    p0 = m.end() - 1
    p1 = _match(src, p0, "(", ")")
    params, depth, cur = [], 0, ""
    for c in src[p0 + 1:p1] + ",":
        if c in "{[(<":
            depth += 1
        elif c in "}])>":
            depth -= 1
        if c == "," and depth == 0:
            if cur.strip():
                params.append(re.match(r"\s*(\w+)", cur).group(1))
            cur = ""
        else:
            cur += c
    b0 = src.index("{", p1)
    b1 = _match(src, b0, "{", "}")
    return params, src[b0 + 1:b1]


# ---- statement splitter -------------------------------------------------------------
def split_statements(body):
    """-> list of tokens: statement strings, '{', '}'.  Statements end at ';' or a newline
    outside (), []; object-literal braces stay inside their statement."""
    out, cur, depth, i, n = [], "", 0, 0, len(body)

    def flush():
        nonlocal cur
        if cur.strip():
            out.append(cur.strip())
        cur = ""

    while i < n:
        c = body[i]
        if c in "([":
            depth += 1
            cur += c
        elif c in ")]":
            depth -= 1
            cur += c
        elif c == "{":
            prev = cur.rstrip()[-1:] if cur.strip() else ""
            if depth > 0 or prev in ("=", ",", "(", ":"):
                j = _match(body, i, "{", "}")      # object literal
                cur += body[i:j + 1]
                i = j
            else:
                flush()
                out.append("{")
        elif c == "}":
            flush()
            out.append("}")
        elif c == ";" and depth == 0:
            flush()
        elif c == "\n" and depth == 0:
            flush()
        elif c in "\"'":
            j = body.index(c, i + 1)
            cur += body[i:j + 1]
            i = j
        else:
            cur += c
        i += 1
    flush()
    return out


# ---- expression rewriting --------------------------------------------------------------
def _object_literals(e):
    while True:
        m = re.search(r"\{([^{}]*)\}", e)
        if not m:
            return e
        inner = m.group(1).strip()
        args = []
        if inner:
            depth, cur = 0, ""
            for c in inner + ",":
                if c in "([":
                    depth += 1
                elif c in ")]":
                    depth -= 1
                if c == "," and depth == 0:
                    k, v = cur.split(":", 1)
                    args.append("%s=%s" % (k.strip(), v.strip()))
                    cur = ""
                else:
                    cur += c
        e = e[:m.start()] + "_obj(%s)" % ", ".join(args) + e[m.end():]


def expr(e):
    e = e.strip()
    e = re.sub(r"\s+as\s+\w+", "", e)                                   # `{} as Config`
    e = _object_literals(e)
    e = re.sub(r"\b(0x[0-9a-fA-F]+|\d+)n\b", r"\1", e)                  # BigInt literals
    e = re.sub(r"\bnew\s+Error\b", "Exception", e)
    e = re.sub(r"\bnew\s+Array\b", "_Array", e)
    e = re.sub(r"\bnew\s+(Float32Array|Int32Array)\b", r"\1", e)
    e = re.sub(r"\bNumber\(", "float(", e)
    e = re.sub(r"\bMath\.", "_Math.", e)
    e = re.sub(r"(\w+)\.length\b", r"len(\1)", e)
    e = re.sub(r"\.charAt\(([^()]*)\)", r"[\1]", e)
    e = re.sub(r"(\w+)\.indexOf\(", r"_indexOf(\1, ", e)
    e = re.sub(r"\(([^?()]+)\?([^:()]+):([^()]+)\)", r"((\2) if (\1) else (\3))", e)   # ternary
    e = re.sub(r"\(([\w\s,]*)\)\s*=>\s*", r"lambda \1: ", e)            # arrow functions
    e = e.replace("===", "==").replace("!==", "!=").replace("&&", " and ").replace("||", " or ")
    e = re.sub(r"\btrue\b", "True", e)
    e = re.sub(r"\bfalse\b", "False", e)
    e = re.sub(r"\bnull\b", "None", e)
    if re.search(r"\+\+|--|=>|\?", e):
        raise SyntaxError("untranslated construct in %r" % e)
    return e


def simple(stmt):
    """One non-control statement -> list of Python lines."""
    s = stmt.strip()
    m = re.match(r"^(let|const|var)\s+(.*)$", s)
    if m:
        s = m.group(2)
        s = re.sub(r"^(\w+)\s*:\s*[\w\[\]<>]+\s*=", r"\1 =", s)         # `let x: T = ...`
    if s == "break":
        return ["break"]
    if s.startswith("return"):
        return ["return " + expr(s[6:])] if s[6:].strip() else ["return"]
    if s.startswith("throw "):
        return ["raise " + expr(s[6:])]
    m = re.match(r"^(\w+)\[(\w+)\+\+\]\s*=\s*(.*)$", s)                 # a[n++] = v
    if m:
        return ["%s[%s] = %s" % (m.group(1), m.group(2), expr(m.group(3))), "%s += 1" % m.group(2)]
    m = re.match(r"^(\+\+|--)?(\w+)(\+\+|--)?$", s)
    if m and (m.group(1) or m.group(3)):
        op = m.group(1) or m.group(3)
        return ["%s %s= 1" % (m.group(2), "+" if op == "++" else "-")]
    m = re.match(r"^([^=(]+?)\s*(\+=|-=|\*=|/=|\^=|=(?![=>]))\s*(.*)$", s)
    if m and not m.group(1).rstrip().endswith(("=", "!", "<", ">")):
        return ["%s %s %s" % (expr(m.group(1)), m.group(2), expr(m.group(3)))]
    return [expr(s)]


def _header(stmt, kw):
    """`kw (...) rest` -> (inside, rest)."""
    i = stmt.index("(")
    j = _match(stmt, i, "(", ")")
    return stmt[i + 1:j], stmt[j + 1:].strip()


def translate(tokens, indent, out, pos=0, single=False):
    """Emit Python for tokens[pos:] until the closing '}' (or one statement if single)."""
    pad = "    " * indent
    emitted = 0
    while pos < len(tokens):
        t = tokens[pos]
        if t == "}":
            if emitted == 0:
                out.append(pad + "pass")
            return pos + 1
        if t == "{":
            raise SyntaxError("unexpected block")
        kw = re.match(r"^(for|while|if|else)\b", t)
        if kw and kw.group(1) == "for":
            inside, rest = _header(t, "for")
            init, cond, upd = [p.strip() for p in inside.split(";")]
            for ln in simple(init):
                out.append(pad + ln)
            out.append(pad + "while %s:" % expr(cond))
            pos = _body(tokens, pos + 1, rest, indent + 1, out)
            for ln in simple(upd):
                out.append(pad + "    " + ln)
        elif kw and kw.group(1) == "while":
            inside, rest = _header(t, "while")
            out.append(pad + "while %s:" % expr(inside))
            pos = _body(tokens, pos + 1, rest, indent + 1, out)
        elif kw and kw.group(1) == "if":
            inside, rest = _header(t, "if")
            out.append(pad + "if %s:" % expr(inside))
            pos = _body(tokens, pos + 1, rest, indent + 1, out)
        elif kw and kw.group(1) == "else":
            out.append(pad + "else:")
            pos = _body(tokens, pos + 1, t[4:].strip(), indent + 1, out)
        else:
            for ln in simple(t):
                out.append(pad + ln)
            pos += 1
        emitted += 1
        if single:
            return pos
    return pos


def _body(tokens, pos, inline_rest, indent, out):
    if inline_rest:                               # `if (c) return i`
        sub = split_statements(inline_rest)
        translate(sub, indent, out, 0, single=True)
        return pos
    if tokens[pos] != "{":
        return translate(tokens, indent, out, pos, single=True)
    return translate(tokens, indent, out, pos + 1)


def translate_function(src, name):
    params, body = extract_function(src, name)
    tokens = split_statements(body) + ["}"]
    out = ["def %s(%s):" % (name, ", ".join(params))]
    assigned = [g for g in GLOBALS_ASSIGNED if re.search(r"\b%s\s*(\^|\+|-|\*|/)?=" % g, body)]
    if assigned:
        out.append("    global " + ", ".join(assigned))
    translate(tokens, 1, out)
    return "\n".join(out)


def load_reference(path=REFERENCE, functions=FUNCTIONS):
    """Namespace with the reference's functions, translated from its source text.
    ns['__python__'] holds the generated Python for inspection."""
    src = _strip_comments(open(path).read())
    ns = {"Float32Array": Float32Array, "Int32Array": Int32Array, "_Array": _Array, "_obj": _obj,
          "_indexOf": _indexOf, "_Math": _Math, "rng_seed": 0, "floatCaster": Float32Array(1)}
    py = []
    for f in functions:
        py.append(translate_function(src, f))
    code = "\n\n".join(py)
    ns["__python__"] = code
    exec(compile(code, "<llama2.ts translated>", "exec"), ns)
    return ns


# ---- building the reference's objects from a checkpoint blob -------------------------
def make_config(hdr):
    """readConfig (llama2.ts:80-93) on the 7 header ints (that function reads a Buffer through
    DataView, which is I/O, not arithmetic; restated here)."""
    c = SimpleNamespace()
    c.dim, c.hidden_dim, c.n_layers, c.n_heads, c.n_kv_heads = [int(v) for v in hdr[:5]]
    c.vocab_size = abs(int(hdr[5]))
    c.seq_len = int(hdr[6])
    c.shared_weights = int(hdr[5]) > 0
    c.head_size = c.dim / c.n_heads
    return c


def make_weights(c, blob):
    """readWeights (llama2.ts:112-129): the same slices, in file order, as Float32Arrays."""
    off = 0

    def get(*dims):
        nonlocal off
        n = int(np.prod(dims))
        a = Float32Array(blob[off:off + n].copy())
        off += n
        return a

    def gets(d0, *dims):
        return [get(*dims) for _ in range(d0)]

    hs2 = int(c.head_size) // 2
    w = SimpleNamespace()
    w.token_embedding_table = get(c.vocab_size, c.dim)
    w.rms_att_weight = gets(c.n_layers, c.dim)
    w.wq = gets(c.n_layers, c.dim, c.dim)
    w.wk = gets(c.n_layers, c.dim, c.dim)
    w.wv = gets(c.n_layers, c.dim, c.dim)
    w.wo = gets(c.n_layers, c.dim, c.dim)
    w.rms_ffn_weight = gets(c.n_layers, c.dim)
    w.w1 = gets(c.n_layers, c.hidden_dim, c.dim)
    w.w2 = gets(c.n_layers, c.dim, c.hidden_dim)
    w.w3 = gets(c.n_layers, c.hidden_dim, c.dim)
    w.rms_final_weight = get(c.dim)
    w.freq_cis_real = get(c.seq_len, hs2)
    w.freq_cis_imag = get(c.seq_len, hs2)
    w.wcls = w.token_embedding_table if c.shared_weights else get(c.vocab_size, c.dim)
    assert off == blob.size
    return w


if __name__ == "__main__":
    print(load_reference()["__python__"])
