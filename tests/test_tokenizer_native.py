"""Native tokenizer: l2b_tok_* vs the golden ids produced by the
reference's own bpe_encode text, vs the oracle restatement, and a synthetic tokenizer.bin so
that the test also runs where the reference checkout is not mounted."""
import os
import struct
import time

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
REF_TOK = "/root/reference/tokenizer.bin"


def synthetic_tokenizer_bin():
    """256 byte-level characters + merges with scores, incl. duplicate strings (first wins)."""
    entries = [(b"<unk>", 0.0), (b"\n<s>\n", 0.0), (b"\n</s>\n", 0.0)]
    entries += [(bytes([c]), -1e3 - c) for c in range(32, 127)]
    merges = ["th", "he", "the", " t", " the", "in", "er", "an", " a", "on", "re", "at", "en", " s", "nd", "and",
              " and", "ing", "ou", "ll", "hello", "he", "el", "lo", "hel", "llo", "wor", "ld", "world", " w", "é", "né"]
    for i, m in enumerate(merges):
        entries.append((m.encode("utf-8"), -float(i)))
    out = struct.pack("<i", max(len(b) for b, _ in entries))
    for b, sc in entries:
        out += struct.pack("<fi", sc, len(b)) + b
    return out, len(entries)


def test_native_matches_oracle_on_synthetic_vocab(pkg, oracle, tmp_path):
    data, n = synthetic_tokenizer_bin()
    path = tmp_path / "tok.bin"
    path.write_bytes(data)
    nat = pkg.Tokenizer(data, n)
    ora = oracle.Tokenizer(str(path), n)
    vocab, scores = pkg.host.read_tokenizer(str(path), n)
    rng = np.random.default_rng(3)
    alphabet = list("the and hello world in on at re er ll ou")
    for trial in range(200):
        s = "".join(rng.choice(alphabet, size=int(rng.integers(1, 60))))
        a = nat.encode(s)
        assert np.array_equal(a, ora.encode(s)), s
        buf = np.zeros(len(s) + 1, np.int32)
        k = pkg.host.bpe_encode(s, vocab, scores, n, buf)
        assert np.array_equal(a, buf[:k]), s
    assert nat.piece(3) == " " and abs(nat.score(3) + 1032.0) < 1e-6
    assert np.array_equal(nat.encode("né"), [n - 1])                  # multi-byte characters are ONE lookup
    with pytest.raises(ValueError):
        nat.encode("tab\tchar")                                         # not in the vocab -> the reference throws
    nat.close()


def test_native_matches_reference_golden_ids(pkg):
    if not os.path.exists(REF_TOK):
        pytest.skip("reference checkout not mounted")
    G = np.load(GOLDEN)
    nat = pkg.Tokenizer(REF_TOK, 32000)
    for s, ids in zip(G["bpe_prompts"], G["bpe_ids"]):
        assert np.array_equal(nat.encode(str(s)), ids[ids >= 0]), s
    assert list(nat.encode("Once upon a time")) == [26222, 2501, 263, 931]
    vocab, _ = pkg.host.read_tokenizer(REF_TOK, 32000)
    for i in (0, 1, 2, 3, 100, 259, 931, 26222, 31999):
        assert nat.piece(i) == vocab[i], i
    # throughput on a long prompt: native hash-map encoder vs the reference's O(n^2 * V) indexOf scan
    from oracle import l2ref
    text = "Once upon a time there was a little girl who loved to play in the garden. " * 6
    t0 = time.perf_counter(); a = nat.encode(text); t_nat = time.perf_counter() - t0
    ora = l2ref.Tokenizer(REF_TOK, 32000)
    t0 = time.perf_counter(); b = ora.encode(text); t_ref = time.perf_counter() - t0
    assert np.array_equal(a, b)
    print("bpe_encode of %d chars: native %.4f s, reference algorithm (C port) %.3f s, %.0fx" %
          (len(text), t_nat, t_ref, t_ref / t_nat))
