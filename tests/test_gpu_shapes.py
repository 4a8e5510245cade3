"""At-shape parity for the BASELINE.json configurations that round 1 only covered with scaled-down
stand-ins: stories42M and stories110M batch-1 (configs[1], configs[2]: teacher-forced logits at 40
positions + a 256-token greedy stream, llama2.ts:465-508) and stories110M with 64 independent
sequences on the tensor-core path (configs[2], sampled sequences against the oracle).
Tolerance (north_star): logits within 1e-4 abs / 1e-3 rel, greedy tokens identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-4, 1e-3


def _make(pkg, arch, seed, std, **kw):
    hdr = pkg.synth.header(arch)
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=std)
    ctx = pkg.Context(hdr, device=0, **kw)
    pkg.synth.upload_blob(ctx, hdr, blob)
    return hdr, blob, ctx


@pytest.mark.parametrize("arch,seed", [("stories42M", 42), ("stories110M", 110)])
def test_batch1_logits_and_greedy_stream_at_shape(pkg, oracle, arch, seed):
    hdr, blob, ctx = _make(pkg, arch, seed, 0.05, max_batch=1, max_steps=256)
    V = abs(hdr[5])
    ref = oracle.Model(hdr, blob)
    oracle.set_threads(oracle.max_threads())
    try:
        toks = np.concatenate([[1], pkg.synth.teacher_tokens(39, V, seed)])
        worst, exact = 0.0, 0
        for pos in range(40):
            got = ctx.forward(int(toks[pos]), pos)
            want = ref.forward(int(toks[pos]), pos)
            assert np.allclose(got, want, rtol=RTOL, atol=ATOL), (pos, np.abs(got - want).max())
            assert int(np.argmax(got)) == oracle.argmax(want)
            worst = max(worst, float(np.abs(got - want).max()))
            exact += int((got == want).sum())
        # `-t 0 -n 256 -i <prompt>`: the device-resident loop and the reference loop, token by token
        prompt = pkg.synth.teacher_tokens(8, V, seed + 1)
        want_stream, _ = oracle.Model(hdr, blob).generate(256, prompt, temperature=0.0)
        ctx.reset()
        forced = np.full(256, -1, dtype=np.int32)
        forced[:8] = prompt
        got_stream = ctx.generate_greedy([1], [0], 256, forced)[:, 0]
        n = len(want_stream)          # the reference stops at BOS (llama2.ts:499)
        assert n > 32 and np.array_equal(got_stream[:n], want_stream)
        print("%s: max|dlogit| %.3g, bit-identical %.3f%%, %d-token greedy stream identical (%d distinct tokens)"
              % (arch, worst, 100.0 * exact / (40 * V), n, len(set(want_stream.tolist()))))
    finally:
        oracle.set_threads(1)
        ctx.close()


def test_stories110m_batch64_tensor_core_path(pkg, oracle):
    """BASELINE configs[2] batched: 64 sequences, 10 teacher-forced steps each; every 8th sequence is
    replayed on the oracle."""
    B, steps = 64, 10
    hdr, blob, ctx = _make(pkg, "stories110M", 64, 0.04, max_batch=B, max_steps=steps)
    V = abs(hdr[5])
    streams = np.stack([np.concatenate([[1], pkg.synth.teacher_tokens(steps - 1, V, 500 + b)]) for b in range(B)])
    sample = list(range(0, B, 8))
    got = np.empty((steps, len(sample), V), dtype=np.float32)
    am = np.empty((steps, B), dtype=np.int32)
    for s in range(steps):
        lg, am[s] = ctx.forward_batch(streams[:, s].astype(np.int32), np.full(B, s, np.int32))
        got[s] = lg[sample]
    assert ctx.last_launches() > 0
    ctx.close()
    oracle.set_threads(oracle.max_threads())
    try:
        worst = 0.0
        for i, b in enumerate(sample):
            ref = oracle.Model(hdr, blob)
            for s in range(steps):
                want = ref.forward(int(streams[b, s]), s)
                err = float(np.abs(got[s, i] - want).max())
                worst = max(worst, err)
                assert np.allclose(got[s, i], want, rtol=RTOL, atol=ATOL), (b, s, err)
                assert am[s, b] == oracle.argmax(got[s, i])
        print("stories110M B=64 (3xTF32 tcgen05): max|dlogit| %.3g over %d sequences x %d steps" % (worst, len(sample), steps))
    finally:
        oracle.set_threads(1)
