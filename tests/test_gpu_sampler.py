"""Device-side temperature / softmax / sample / sample_topp against the
oracle's restatement of llama2.ts:476-494, 368-394: same token for the same random number."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch", ["tiny", "stories15M"])
def test_sampler_on_given_logits(pkg, oracle, arch):
    hdr = pkg.synth.header(arch)
    V = abs(hdr[5])
    rng = np.random.default_rng(7)
    n_bad = 0
    with pkg.Context(hdr, max_steps=4) as ctx:
        trials = 0
        for trial in range(60):
            kind = trial % 6
            if kind == 0:
                logits = (rng.standard_normal(V) * 6).astype(np.float32)           # peaked
            elif kind == 1:
                logits = (rng.standard_normal(V) * 0.05).astype(np.float32)        # nearly uniform
            elif kind == 2:
                logits = np.zeros(V, np.float32)                                   # exactly uniform
            elif kind == 3:
                logits = (rng.standard_normal(V) * 3).astype(np.float32)
                logits[rng.integers(0, V, 5)] = logits.max()                       # ties at the top
            elif kind == 4:
                logits = np.round(rng.standard_normal(V) * 2).astype(np.float32)   # many exact ties
            else:
                logits = (rng.standard_normal(V) * 2).astype(np.float32)
            for temperature, topp in ((1.0, 1.0), (0.7, 0.9), (1.5, 0.5), (1.0, 0.05), (0.3, 0.999), (1.0, 0.0)):
                seed = int(rng.integers(1, 2**31))
                want = oracle.sample_next(logits, temperature, topp, oracle.Rng(seed))
                r = oracle.Rng(seed).f32()
                got = ctx.sample_logits(logits, temperature, topp, r)
                trials += 1
                if got != want:
                    n_bad += 1
                    print("mismatch", arch, kind, temperature, topp, seed, got, want)
        assert n_bad == 0, "%d of %d sampler choices differ" % (n_bad, trials)


@pytest.mark.parametrize("temperature,topp", [(1.0, 1.0), (0.8, 0.9), (1.2, 0.5)])
def test_generate_loop_with_device_sampler(pkg, oracle, temperature, topp):
    """The host loop with forward+sampler fused into one call emits the reference's tokens."""
    hdr = pkg.synth.header("small")
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=71, std=0.08)
    H = pkg.host
    prompt = np.array([17, 300, 45], dtype=np.int32)
    import struct
    config = H.readConfig(struct.pack("<7i", *hdr))
    for seed in (1, 99):
        want, _ = oracle.Model(hdr, blob).generate(80, prompt, temperature=temperature, topp=topp, seed=seed)
        ctx = pkg.Context(hdr, max_steps=80)
        pkg.synth.upload_blob(ctx, hdr, blob)
        got, _ = H.generate(config, H.TransformerWeights(ctx), H.newRunState(config), 80, prompt, temperature,
                            topp, H.Rng(seed), device_sampler=True)
        assert np.array_equal(np.array(got), want), (temperature, topp, seed)
        ctx.close()
