"""Checkpoint loader fast path: l2b_load_checkpoint must leave the
device in exactly the state the reference-order l2b_upload calls do."""
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch", ["tiny-unshared", "small", "stories15M"])
def test_file_loader_equals_per_tensor_upload(pkg, oracle, arch, tmp_path):
    hdr = pkg.synth.header(arch)
    path = pkg.synth.write_checkpoint(str(tmp_path / "m.bin"), hdr, seed=61, std=0.05)
    blob = np.fromfile(path, dtype=np.float32, offset=28)
    V = abs(hdr[5])
    toks = np.concatenate([[1], pkg.synth.teacher_tokens(9, V, 61)])
    with pkg.Context(hdr, max_steps=16) as a, pkg.Context(hdr, max_steps=16) as b:
        pkg.synth.upload_blob(a, hdr, blob)
        assert not b.weights_ready()
        secs = b.load_checkpoint(path)
        assert b.weights_ready() and secs > 0
        for pos, t in enumerate(toks):
            assert np.array_equal(a.forward(int(t), pos), b.forward(int(t), pos)), pos
        want = oracle.Model(hdr, blob).forward(1, 0)
        b.reset()
        assert np.allclose(b.forward(1, 0), want, rtol=1e-3, atol=1e-4)
    print("%s: %.1f MB in %.3f s" % (arch, os.path.getsize(path) / 1e6, secs))


def test_loader_rejects_bad_files(pkg, tmp_path):
    hdr = pkg.synth.header("tiny")
    good = pkg.synth.write_checkpoint(str(tmp_path / "g.bin"), hdr, seed=1)
    raw = open(good, "rb").read()
    other = str(tmp_path / "o.bin")
    open(other, "wb").write(struct.pack("<7i", 64, 176, 2, 4, 4, 512, 64) + raw[28:])   # seq_len differs
    short = str(tmp_path / "s.bin")
    open(short, "wb").write(raw[:len(raw) // 2])
    with pkg.Context(hdr, max_steps=8) as ctx:
        for bad in (other, short, str(tmp_path / "missing.bin")):
            with pytest.raises(pkg.L2BError) as e:
                ctx.load_checkpoint(bad)
            assert e.value.code == pkg.capi.EINVAL
        assert not ctx.weights_ready()
        ctx.load_checkpoint(good)
        assert ctx.weights_ready()
