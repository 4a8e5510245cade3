"""Host-side logic (no GPU): checkpoint format, tokenizer, RNG and samplers of the host
mirror against the oracle's restatement of the same reference lines."""
import os
import struct

import numpy as np
import pytest

TOKENIZER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tokenizer_head.bin")


def test_named_architectures_match_survey_file_sizes(pkg):
    # file bytes of the llama2.c checkpoints
    want = {"stories15M": 60816028, "stories42M": 167020572, "stories110M": 438381596,
            "llama2-7b": 26954711068}
    for name, nbytes in want.items():
        hdr = pkg.synth.header(name)
        assert 28 + 4 * pkg.synth.weight_floats(hdr) == nbytes, name
    assert pkg.synth.header("llama2-7b")[5] == -32000           # unshared classifier
    # SURVEY.md 8(d) bytes/token table
    assert abs(pkg.synth.step_bytes(pkg.synth.header("stories15M"), 0) / 1e6 - 60.80) < 0.01
    assert abs(pkg.synth.step_bytes(pkg.synth.header("llama2-7b"), 255) / 1e6 - 26699) < 1


def test_checkpoint_file_roundtrip(pkg, oracle, tmp_path):
    hdr = pkg.synth.header("tiny-unshared")
    path = pkg.synth.write_checkpoint(str(tmp_path / "t.bin"), hdr, seed=5)
    raw = open(path, "rb").read()
    assert list(struct.unpack("<7i", raw[:28])) == hdr
    c = pkg.host.readConfig(raw[:28])
    assert (c.dim, c.hidden_dim, c.n_layers, c.n_heads, c.vocab_size, c.seq_len) == (96, 256, 3, 2, 1000, 48)
    assert c.shared_weights is False and c.head_size == 48
    blob = np.frombuffer(raw, dtype=np.float32, offset=28)
    _, blob2 = pkg.synth.checkpoint_blob(hdr, seed=5)
    assert np.array_equal(blob, blob2)
    m = oracle.Model(hdr, blob)                                  # the oracle accepts the same file
    assert np.isfinite(m.forward(1, 0)).all()


def test_rng_stream_matches_reference_kat(pkg, oracle):
    # SURVEY.md 8(c) KAT for `-s 1` (llama2.ts:349-360)
    r = pkg.host.Rng(1)
    assert [r.random_u32() for _ in range(4)] == [1206177355, 2882512552, 3117485455, 1303648416]
    r, o = pkg.host.Rng(12345), oracle.Rng(12345)
    for _ in range(1000):
        assert r.random_f32() == o.f32()


@pytest.mark.parametrize("temperature,topp", [(0.0, 1.0), (1.0, 1.0), (0.7, 1.0), (1.0, 0.9), (0.5, 0.5),
                                              (1.3, 0.99), (1.0, 0.0)])
def test_samplers_match_oracle(pkg, oracle, temperature, topp):
    """Host mirror samplers (llama2.ts:364-394, 476-494) vs the oracle restatement on random
    logits: identical token for identical seed, including the top-p quirks."""
    H = pkg.host
    rng = np.random.default_rng(0)
    for trial in range(40):
        V = int(rng.integers(8, 3000))
        logits = (rng.standard_normal(V) * rng.uniform(0.5, 6)).astype(np.float32)
        if trial % 5 == 0:
            logits[rng.integers(0, V, 3)] = logits.max()          # ties
        seed = int(rng.integers(1, 2**31))
        want = oracle.sample_next(logits, temperature, topp, oracle.Rng(seed))
        r = H.Rng(seed)
        lg = logits.copy()
        if temperature == 0.0:
            got = H.argmax(lg)
        else:
            lg[:] = (lg.astype(np.float64) / temperature).astype(np.float32)
            H.softmax(lg, 0, V)
            got = H.sample(lg, V, r) if (topp <= 0 or topp >= 1) else H.sample_topp(lg, topp, r)
        assert got == want, (trial, V, seed)


def test_argmax_nan_and_ties(pkg, oracle):
    H = pkg.host
    a = np.array([1, 5, 5, 2], dtype=np.float32)
    assert H.argmax(a) == oracle.argmax(a) == 1
    a = np.array([np.nan, 5, 7], dtype=np.float32)
    assert H.argmax(a) == oracle.argmax(a) == 0                   # NaN at 0 is never beaten
    a = np.array([1, np.nan, 7, np.nan], dtype=np.float32)
    assert H.argmax(a) == oracle.argmax(a) == 2
