"""Single-process multi-GPU contexts (l2b_create_multi): ONE host thread -- like the
reference's single JS thread at llama2.ts:468 -- drives every device.  The 1-GPU cases run on any box
and exercise the group dispatch of every entry point; the 2-GPU cases need `gpurun --gpus 2`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-4, 1e-3


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _model(pkg, arch, seed, std=0.05):
    hdr = pkg.synth.header(arch)
    _, blob = pkg.synth.checkpoint_blob(hdr, seed=seed, std=std)
    return hdr, blob


def test_group_of_one_is_bit_identical_to_plain_context(pkg, oracle):
    hdr, blob = _model(pkg, "small", 61)
    V, S = abs(hdr[5]), 24
    toks = np.concatenate([[1], pkg.synth.teacher_tokens(S - 1, V, 61)])
    with pkg.Context(hdr, n_gpus=1, tp_degree=1, max_batch=1, max_steps=S) as grp, \
            pkg.Context(hdr, device=0, max_batch=1, max_steps=S) as one:
        pkg.synth.upload_blob(grp, hdr, blob)
        pkg.synth.upload_blob(one, hdr, blob)
        assert grp.weights_ready()
        for pos in range(S):
            assert np.array_equal(grp.forward(int(toks[pos]), pos), one.forward(int(toks[pos]), pos))
            assert grp.forward_argmax(int(toks[pos]), pos) == one.forward_argmax(int(toks[pos]), pos)
        assert np.array_equal(grp.read_state(pkg.capi.S_KEY_ROW, 0, 1, 3), one.read_state(pkg.capi.S_KEY_ROW, 0, 1, 3))
        grp.reset(); one.reset()
        forced = np.full(S, -1, np.int32); forced[:3] = toks[1:4]
        a = grp.generate_greedy([1], [0], S, forced)
        b = one.generate_greedy([1], [0], S, forced)
        assert np.array_equal(a, b) and grp.last_launches() == one.last_launches() and grp.last_device_ms() > 0
        # device sampler and the order check go through the group as well
        grp.reset(); one.reset()
        assert grp.forward_sample(1, 0, 0.8, 0.9, 0.37) == one.forward_sample(1, 0, 0.8, 0.9, 0.37)
        with pytest.raises(pkg.capi.L2BError) as e:
            grp.forward(5, 7)
        assert e.value.code == pkg.capi.EORDER and "pos 7" in str(e.value)


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_batch_partition_group_vs_oracle(pkg, oracle, n_gpus):
    """tp_degree 1: sequence b lives on GPU b / ceil(max_batch / n_gpus); weights replicated."""
    if _ngpu() < n_gpus:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (n_gpus, n_gpus))
    hdr, blob = _model(pkg, "small", 62, 0.08)
    V, B, steps = abs(hdr[5]), 7, 12
    first = pkg.synth.teacher_tokens(B, V, 62)
    forced = np.full((steps, B), -1, np.int32)
    forced[0] = first
    oracle.set_threads(oracle.max_threads())
    want = [oracle.Model(hdr, blob).generate(steps, [int(first[b])], temperature=0.0)[0] for b in range(B)]
    refs = [oracle.Model(hdr, blob) for _ in range(B)]
    with pkg.Context(hdr, n_gpus=n_gpus, tp_degree=1, max_batch=B, max_steps=steps) as grp:
        pkg.synth.upload_blob(grp, hdr, blob)
        got = grp.generate_greedy(np.ones(B, np.int32), np.zeros(B, np.int32), steps, forced)
        for b in range(B):
            assert np.array_equal(got[:len(want[b]), b], want[b]), b
        grp.reset()
        for s in range(3):     # logits of every sequence, gathered from the members
            toks = np.ones(B, np.int32) if s == 0 else forced[0] if s == 1 else np.full(B, 9, np.int32)
            lg, am = grp.forward_batch(toks.astype(np.int32), np.full(B, s, np.int32))
            for b in range(B):
                w = refs[b].forward(int(toks[b]), s)
                assert np.allclose(lg[b], w, rtol=RTOL, atol=ATOL), (s, b, np.abs(lg[b] - w).max())
                assert am[b] == oracle.argmax(lg[b])
        # a call with fewer sequences than max_batch uses the first members only
        lg, _ = grp.forward_batch(np.full(2, 9, np.int32), np.full(2, 3, np.int32))
        assert lg.shape == (2, V)
    oracle.set_threads(1)


@pytest.mark.parametrize("arch,steps", [("small", 24), ("wide", 20)])
def test_single_process_tensor_parallel(pkg, oracle, arch, steps):
    """tp_degree == n_gpus == 2 in ONE process: logits bit-identical to the one-GPU library (row
    sharding keeps every output's summation order), inside the tolerance of the oracle."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    hdr, blob = _model(pkg, arch, 63, 0.04)
    V = abs(hdr[5])
    toks = np.concatenate([[1], pkg.synth.teacher_tokens(steps - 1, V, 63)])
    ref = oracle.Model(hdr, blob)
    oracle.set_threads(4)
    with pkg.Context(hdr, n_gpus=2, tp_degree=2, max_batch=1, max_steps=steps) as tp, \
            pkg.Context(hdr, device=0, max_batch=1, max_steps=steps) as one:
        pkg.synth.upload_blob(tp, hdr, blob)
        pkg.synth.upload_blob(one, hdr, blob)
        one.set_option("fuse_qkv_attn", 0)     # a tensor-parallel rank runs the stand-alone q/k/v and attention kernels
        for pos in range(steps):
            got = tp.forward(int(toks[pos]), pos)
            want = ref.forward(int(toks[pos]), pos)
            assert np.allclose(got, want, rtol=RTOL, atol=ATOL), (pos, np.abs(got - want).max())
            assert np.array_equal(got, one.forward(int(toks[pos]), pos)), pos
        tp.reset(); one.reset()
        forced = np.full(steps, -1, np.int32); forced[:3] = toks[1:4]
        a = tp.generate_greedy([1], [0], steps, forced)[:, 0]
        b = one.generate_greedy([1], [0], steps, forced)[:, 0]
        assert np.array_equal(a, b)
        tp.reset()
        assert tp.forward_sample(1, 0, 0.8, 0.9, 0.37) == one.forward_sample(1, 0, 0.8, 0.9, 0.37)
    oracle.set_threads(1)


def test_group_upload_from_device_memory_keeps_callers_device(pkg, oracle):
    """Weights handed over as DEVICE pointers on GPU 0 (l2b_upload copies with cudaMemcpyDefault; the
    members on other GPUs copy across NVLink), generated tensor by tensor like bench.py does -- and the
    calling thread's current device is restored by every entry point (a group walks over all GPUs)."""
    import torch
    n = 2 if _ngpu() >= 2 else 1
    hdr, blob = _model(pkg, "wide", 64, 0.04)
    V, steps = abs(hdr[5]), 8
    toks = np.concatenate([[1], pkg.synth.teacher_tokens(steps - 1, V, 64)])
    torch.cuda.set_device(0)
    with pkg.Context(hdr, n_gpus=n, tp_degree=n, max_batch=1, max_steps=steps) as grp, \
            pkg.Context(hdr, device=0, max_batch=1, max_steps=steps) as one:
        for (t, l), a in pkg.synth.slice_blob(hdr, blob).items():
            ta = torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0") * 1.0   # a kernel produces the tensor
            torch.cuda.synchronize()                 # synchronises the CURRENT device: must still be 0
            grp.upload(t, l, ta)
            assert torch.cuda.current_device() == 0
            one.upload(t, l, ta)
            del ta
        one.set_option("fuse_qkv_attn", 0 if n > 1 else 1)
        for pos in range(steps):
            a = grp.forward(int(toks[pos]), pos)
            assert torch.cuda.current_device() == 0
            assert np.array_equal(a, one.forward(int(toks[pos]), pos)), pos
