"""Pins oracle/l2ref.c against golden vectors produced by executing the reference's own
source text (tests/golden/make_golden.py via oracle/ts_exec.py).  Bit-exact: the oracle is a
restatement of the same f64-temporaries / f32-stores arithmetic in the same order."""
import hashlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


def same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


@pytest.mark.parametrize("n", [1, 2, 7, 48, 288])
def test_primitives_bit_exact(oracle, G, n):
    x, w = G["rmsnorm_%d_x" % n], G["rmsnorm_%d_w" % n]
    assert same(oracle.rmsnorm(x, w), G["rmsnorm_%d_o" % n])                 # llama2.ts:172-179
    assert same(oracle.softmax(G["softmax_%d_x" % n]), G["softmax_%d_o" % n])  # :181-194
    assert same(oracle.accum(x, w), G["accum_%d_o" % n])                     # :168-170


@pytest.mark.parametrize("d,n", [(3, 5), (16, 64), (10, 288)])
def test_matmul_bit_exact(oracle, G, d, n):
    k = "matmul_%dx%d" % (d, n)
    assert same(oracle.matmul(G[k + "_x"], G[k + "_w"]), G[k + "_o"])         # :196-203
    oracle.set_threads(4)                                                    # row split: same bits
    assert same(oracle.matmul(G[k + "_x"], G[k + "_w"]), G[k + "_o"])
    oracle.set_threads(1)


def test_hand_computed_kats(oracle):
    """Tiny cases worked by hand from the reference's formulas."""
    # rmsnorm: x=[3,4], w=[1,2]: ss=(9+16)/2=12.5, s=1/sqrt(12.50001)
    s = 1.0 / np.sqrt(1e-5 + 12.5)
    got = oracle.rmsnorm(np.array([3, 4], np.float32), np.array([1, 2], np.float32))
    assert same(got, np.array([1 * (s * 3), 2 * (s * 4)], np.float64).astype(np.float32))
    # softmax of equal values is uniform; of [0, ln 3] is [0.25, 0.75]
    assert same(oracle.softmax(np.zeros(4, np.float32)), np.full(4, 0.25, np.float32))
    p = oracle.softmax(np.array([0.0, np.log(3.0)], np.float32))
    assert abs(p[0] - 0.25) < 1e-7 and abs(p[1] - 0.75) < 1e-7
    # RoPE: rotating (1,0) by angle a gives (cos a, sin a); same angle for every head
    hs, pos = 4, 3
    ang = np.array([[0.0, 0.0], [0.1, 0.01], [0.2, 0.02], [0.3, 0.03]], np.float32)
    fcr, fci = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    q = np.array([1, 0, 0, 1, 1, 0, 0, 1], np.float32)
    q2, k2 = oracle.rope(q, q.copy(), fcr.ravel(), fci.ravel(), pos, hs)
    want = np.array([fcr[3, 0], fci[3, 0], -fci[3, 1], fcr[3, 1]] * 2, np.float32)
    assert same(q2, want) and same(k2, want)
    # matmul accumulates in f64: 1e8 + 1 - 1e8 survives, where f32 accumulation gives 0
    W = np.array([[1e8, 1.0, -1e8]], np.float32)
    assert oracle.matmul(np.ones(3, np.float32), W)[0] == 1.0


def test_rng_and_samplers_match_reference(oracle, G, pkg):
    for seed in (1, 42, 2**40 + 12345):
        r = oracle.Rng(seed)
        assert [r.u32() for _ in range(16)] == [int(v) for v in G["rng_u32_%d" % seed]]
        r = oracle.Rng(seed)
        assert same(np.array([r.f32() for _ in range(16)], np.float32), G["rng_f32_%d" % seed])
        h = pkg.host.Rng(seed)                                               # host mirror too
        assert same(np.array([h.random_f32() for _ in range(16)], np.float32), G["rng_f32_%d" % seed])
    import ctypes as C
    L = oracle.lib()
    for probs, row in zip(G["sampler_probs"], G["sampler_rows"]):
        V, seed, s_plain, t9, t5, t1, am = [int(v) for v in row]
        p = np.ascontiguousarray(probs[:V])
        pp = p.ctypes.data_as(C.POINTER(C.c_float))
        assert L.l2ref_sample(pp, V, C.byref(C.c_uint64(seed))) == s_plain          # :368-376
        for topp, want in ((0.9, t9), (0.5, t5), (0.1, t1)):                        # :378-394
            assert L.l2ref_sample_topp(pp, V, topp, C.byref(C.c_uint64(seed))) == want
            assert pkg.host.sample_topp(p, topp, pkg.host.Rng(seed)) == want
        assert oracle.argmax(p) == am                                               # :364-366
        assert pkg.host.argmax(p) == am
        assert pkg.host.sample(p, V, pkg.host.Rng(seed)) == s_plain


@pytest.mark.parametrize("key,arch", [("tiny_s1", "tiny"), ("tiny_s21", "tiny"),
                                      ("tiny_unshared_s2", "tiny-unshared")])
def test_transformer_forward_bit_exact(oracle, G, pkg, key, arch):
    """transformer() (llama2.ts:205-303): logits of every step, final x, KV cache."""
    hdr = [int(v) for v in G[key + "_hdr"]]
    seed = int(key.rsplit("s", 1)[1])
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=float(G[key + "_std"][0]))
    assert hashlib.sha256(blob.tobytes()).digest() == G[key + "_sha256"].tobytes(), \
        "synthetic checkpoint drifted: re-run tests/golden/make_golden.py"
    m = oracle.Model(hdr, blob)
    toks = G[key + "_tokens"]
    for pos, t in enumerate(toks):
        assert same(m.forward(int(t), pos), G[key + "_logits"][pos]), pos
    assert same(m.x(), G[key + "_x"])
    for l in range(hdr[2]):
        for pos in range(len(toks)):
            assert same(m.key_row(l, pos), G[key + "_key_cache"][l, pos])
            assert same(m.value_row(l, pos), G[key + "_value_cache"][l, pos])


GOLDEN2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz")


@pytest.mark.parametrize("key,arch", [("small_s5", "small"), ("wide_s6", "wide"), ("stories15M_s7", "stories15M")])
def test_transformer_forward_bit_exact_at_kernel_shapes(oracle, pkg, key, arch):
    """The same pin at the shapes the kernels are specialised for (golden_v2.npz, produced by executing the
    reference's text): head_size 32 / 128 / 48, multi-tile rows, unshared and shared classifier, vocab 32000."""
    G2 = np.load(GOLDEN2)
    hdr = [int(v) for v in G2[key + "_hdr"]]
    assert hdr == pkg.synth.header(arch)
    seed = int(key.rsplit("s", 1)[1])
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=float(G2[key + "_std"][0]))
    assert hashlib.sha256(blob.tobytes()).digest() == G2[key + "_sha256"].tobytes(), \
        "synthetic checkpoint drifted: re-run tests/golden/make_golden.py --v2"
    m = oracle.Model(hdr, blob)
    oracle.set_threads(4)
    toks = G2[key + "_tokens"]
    for pos, t in enumerate(toks):
        assert same(m.forward(int(t), pos), G2[key + "_logits"][pos]), pos
    oracle.set_threads(1)
    assert same(m.x(), G2[key + "_x"])
    for l in range(hdr[2]):
        for pos in range(len(toks)):
            assert same(m.key_row(l, pos), G2[key + "_key_cache"][l, pos])
            assert same(m.value_row(l, pos), G2[key + "_value_cache"][l, pos])


@pytest.mark.parametrize("name,temp,topp", [("greedy", 0.0, 1.0), ("sample", 1.0, 1.0), ("topp", 0.8, 0.9)])
def test_generate_loop_matches_reference(oracle, G, pkg, name, temp, topp):
    """The generate loop (llama2.ts:460-508) incl. prompt forcing, temperature, top-p, seed 1."""
    key = "tiny_s1"
    hdr = [int(v) for v in G[key + "_hdr"]]
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=1, std=float(G[key + "_std"][0]))
    m = oracle.Model(hdr, blob)
    out, lg = m.generate(14, G[key + "_gen_prompt"], temperature=temp, topp=topp, seed=1,
                         want_logits=(name == "greedy"))
    assert np.array_equal(out, G["%s_gen_%s_tokens" % (key, name)])
    if lg is not None:
        assert same(lg, G["%s_gen_%s_logits" % (key, name)])


def test_tokenizer_matches_reference(oracle, G, pkg):
    """bpe_encode (llama2.ts:305-344) on the reference's tokenizer.bin (only where the
    reference checkout is mounted; the GPU box does not have it)."""
    tk = "/root/reference/tokenizer.bin"
    if not os.path.exists(tk):
        pytest.skip("reference checkout not mounted")
    assert hashlib.sha256(open(tk, "rb").read()).digest() == G["tokenizer_sha256"].tobytes()
    tok = oracle.Tokenizer(tk)
    vocab, scores = pkg.host.read_tokenizer(tk, 32000)
    for s, ids in zip(G["bpe_prompts"], G["bpe_ids"]):
        want = ids[ids >= 0]
        assert np.array_equal(tok.encode(str(s)), want), s
        buf = np.zeros(len(str(s)) + 1, dtype=np.int32)
        n = pkg.host.bpe_encode(str(s), vocab, scores, 32000, buf)
        assert np.array_equal(buf[:n], want), s
    assert list(tok.encode("Once upon a time")) == [26222, 2501, 263, 931]
