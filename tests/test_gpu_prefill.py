"""Prompt prefill as one batched tensor-core pass vs the reference's
token-by-token prompt loop (llama2.ts:465-474): KV rows, last logits, and the decode steps
that follow must match the oracle within the 1e-4 / 1e-3 tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-4, 1e-3


@pytest.mark.parametrize("arch,n,std", [("small", 37, 0.05), ("small", 100, 0.05), ("wide", 50, 0.03),
                                        ("stories15M", 200, 0.03), ("stories42M", 300, 0.02)])
def test_prefill_matches_token_by_token(pkg, oracle, arch, n, std):
    hdr = pkg.synth.header(arch)
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=51, std=std)
    V, L = abs(hdr[5]), hdr[2]
    steps = min(hdr[6], n + 6)
    toks = np.concatenate([[1], pkg.synth.teacher_tokens(steps - 1, V, 51)]).astype(np.int32)
    ref = oracle.Model(hdr, blob)
    oracle.set_threads(oracle.max_threads())
    want = [ref.forward(int(toks[p]), p) for p in range(steps)]
    with pkg.Context(hdr, max_steps=steps) as ctx:
        pkg.synth.upload_blob(ctx, hdr, blob)
        lg, am = ctx.prefill(toks[:n], 0)
        assert np.allclose(lg, want[n - 1], rtol=RTOL, atol=ATOL), np.abs(lg - want[n - 1]).max()
        assert am == oracle.argmax(lg)
        for l in (0, L - 1):
            for p in (0, n // 2, n - 1):
                k = ctx.read_state(pkg.capi.S_KEY_ROW, 0, l, p)
                v = ctx.read_state(pkg.capi.S_VALUE_ROW, 0, l, p)
                assert np.allclose(k, ref.key_row(l, p), rtol=RTOL, atol=ATOL), (l, p)
                assert np.allclose(v, ref.value_row(l, p), rtol=RTOL, atol=ATOL), (l, p)
        worst = float(np.abs(lg - want[n - 1]).max())
        for p in range(n, steps):                      # decoding continues on the prefilled cache
            got = ctx.forward(int(toks[p]), p)
            assert np.allclose(got, want[p], rtol=RTOL, atol=ATOL), (p, np.abs(got - want[p]).max())
            worst = max(worst, float(np.abs(got - want[p]).max()))
        # a second prompt chunk appended later (pos0 > 0) and the order check
        ctx.reset()
        ctx.prefill(toks[:n // 2], 0)
        lg2, _ = ctx.prefill(toks[n // 2:n], n // 2)
        assert np.allclose(lg2, want[n - 1], rtol=RTOL, atol=ATOL)
        with pytest.raises(pkg.L2BError) as e:
            ctx.prefill(toks[:2], n + 3)
        assert e.value.code in (pkg.capi.EORDER, pkg.capi.EINVAL)
    oracle.set_threads(1)
    print("%s prefill %d tokens: max|dlogit| %.3g" % (arch, n, worst))


def test_host_loop_with_prefill_emits_same_tokens(pkg, oracle):
    """generate(..., prefill=True): the prompt goes through l2b_prefill, sampling is unchanged."""
    hdr = pkg.synth.header("small")
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=52, std=0.08)
    H = pkg.host
    prompt = pkg.synth.teacher_tokens(20, abs(hdr[5]), 9)
    for temperature, topp in ((0.0, 1.0), (0.9, 0.9)):
        want, _ = oracle.Model(hdr, blob).generate(60, prompt, temperature=temperature, topp=topp, seed=7)
        config = H.readConfig(__import__("struct").pack("<7i", *hdr))
        ctx = pkg.Context(hdr, max_steps=60)
        pkg.synth.upload_blob(ctx, hdr, blob)
        got, _ = H.generate(config, H.TransformerWeights(ctx), H.newRunState(config), 60, prompt, temperature,
                            topp, H.Rng(7), prefill=True)
        assert np.array_equal(np.array(got), want), (temperature, topp)
        ctx.close()
