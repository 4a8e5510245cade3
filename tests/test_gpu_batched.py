"""Batched decode (B independent sequences) on the tcgen05 3xTF32 GEMM path vs the oracle,
sequence by sequence.  Tolerance: 1e-4 abs / 1e-3 rel on every logit (north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-4, 1e-3


def run_batch(pkg, oracle, arch, B, steps, seed, std, opts=()):
    hdr = pkg.synth.header(arch)
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=std)
    V = abs(hdr[5])
    ctx = pkg.Context(hdr, max_batch=B, max_steps=steps)
    pkg.synth.upload_blob(ctx, hdr, blob)
    for k, v in opts:
        ctx.set_option(k, v)
    oracle.set_threads(oracle.max_threads())
    refs = [oracle.Model(hdr, blob) for _ in range(B)]
    streams = [np.concatenate([[1], pkg.synth.teacher_tokens(steps - 1, V, 1000 + b)]) for b in range(B)]
    pos = np.zeros(B, dtype=np.int32)
    worst = 0.0
    for step in range(steps):
        toks = np.array([streams[b][pos[b]] for b in range(B)], dtype=np.int32)
        logits, am = ctx.forward_batch(toks, pos)
        for b in range(B):
            want = refs[b].forward(int(toks[b]), int(pos[b]))
            err = np.abs(logits[b] - want).max()
            worst = max(worst, float(err))
            assert np.allclose(logits[b], want, rtol=RTOL, atol=ATOL), (step, b, err)
            assert am[b] == oracle.argmax(logits[b])
        for b in range(B):
            if step >= b % 4:          # staggered positions across the batch
                pos[b] += 1
    oracle.set_threads(1)
    launches = ctx.last_launches()
    ctx.close()
    return worst, launches


@pytest.mark.parametrize("arch,B,std", [("tiny", 9, 0.05), ("tiny", 40, 0.05), ("small", 16, 0.05),
                                        ("small", 70, 0.03), ("wide", 24, 0.03), ("wide", 130, 0.03)])
def test_tensor_core_path_vs_oracle(pkg, oracle, arch, B, std):
    worst, launches = run_batch(pkg, oracle, arch, B, 10, 17, std)
    print("%s B=%d: max|dlogit| %.3g (%d launches/step)" % (arch, B, worst, launches))


@pytest.mark.parametrize("splits", [1, 2, 3, 4])
def test_k_splits(pkg, oracle, splits):
    worst, _ = run_batch(pkg, oracle, "wide", 20, 6, 23, 0.03, opts=(("tc_splits", splits),))
    print("splits=%d: max|dlogit| %.3g" % (splits, worst))


@pytest.mark.parametrize("B", [20, 200])
def test_weight_hi_rounding_variant(pkg, oracle, B):
    """tc_rewrite_hi=1: the weights' hi part is rounded to nearest and rewritten in shared
    memory instead of relying on the tensor core truncating the raw fp32 tile."""
    worst, _ = run_batch(pkg, oracle, "wide", B, 5, 23, 0.03, opts=(("tc_rewrite_hi", 1),))
    print("rewrite_hi B=%d: max|dlogit| %.3g" % (B, worst))


def test_batched_greedy_loop_matches_gemv_path(pkg, oracle):
    """Device-resident greedy loop for B sequences: tensor-core path and the fp64 GEMV path
    (tc_min_batch=0) must emit the same token streams as B separate reference loops."""
    hdr = pkg.synth.header("small")
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=29, std=0.08)
    B, steps, V = 12, 40, abs(hdr[5])
    first = pkg.synth.teacher_tokens(B, V, 5)
    forced = np.full((steps, B), -1, dtype=np.int32)
    forced[0] = first                      # a different first real token per sequence
    want = []
    oracle.set_threads(oracle.max_threads())
    for b in range(B):
        m = oracle.Model(hdr, blob)
        out, _ = m.generate(steps, [int(first[b])], temperature=0.0)
        want.append(out)
    oracle.set_threads(1)
    for tc in (5, 0):
        ctx = pkg.Context(hdr, max_batch=B, max_steps=steps)
        pkg.synth.upload_blob(ctx, hdr, blob)
        ctx.set_option("tc_min_batch", tc)
        got = ctx.generate_greedy(np.ones(B, np.int32), np.zeros(B, np.int32), steps, forced)
        for b in range(B):
            n = len(want[b])
            assert np.array_equal(got[:n, b], want[b]), (tc, b)
        ctx.close()
