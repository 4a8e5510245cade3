"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports
every symbol include/llama2_b200.h declares, and refuses to run without a GPU (no CPU path)."""
import ctypes as C
import os
import subprocess

import pytest


def test_library_builds_and_exports_every_declared_symbol(pkg):
    lib = pkg.capi.Library.get()
    names = pkg.capi.declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert lib.exported(n), n
        assert n in pkg.capi._SIGNATURES, "capi.py has no prototype for %s" % n
    assert lib.dll.l2b_abi_version() == 2


def test_library_is_sm100a_and_has_no_torch_dependency(pkg):
    so = pkg.capi.LIB_PATH
    elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "libc10" not in ldd


def test_sass_uses_bulk_copy_and_fp64(pkg):
    """Evidence that the attention kernel stages the KV cache with TMA bulk copies (UBLKCP)
    and that the matvec accumulates in f64 (DFMA)."""
    sass = subprocess.run(["cuobjdump", "-sass", pkg.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    assert "DFMA" in sass
    assert "LDG.E.128" in sass


def test_sass_uses_tcgen05_and_tensor_memory(pkg):
    """Evidence that the batched path runs on the 5th-generation tensor cores: tcgen05.mma (UTCHMMA) with
    accumulators / the weight operand in tensor memory (LDTM reads, STTM writes), no mma.sync / wgmma recompiles."""
    sass = subprocess.run(["cuobjdump", "-sass", pkg.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert sass.count("UTCHMMA") >= 16
    assert "LDTM" in sass and "STTM" in sass
    assert "HMMA.16816" not in sass and "WGMMA" not in sass


def test_multi_gpu_argument_validation_needs_no_gpu(pkg):
    """l2b_create_multi (SURVEY 8b's signature): bad n_gpus / tp_degree / max_batch are rejected before any
    device is touched; group-only restrictions are stated in the message."""
    lib = pkg.capi.Library.get()
    out = C.c_void_p()
    hdr = (C.c_int32 * 7)(64, 176, 2, 4, 4, 512, 32)
    assert lib.dll.l2b_create_multi(hdr, 0, 1, 1, 0, C.byref(out)) == pkg.capi.EINVAL     # n_gpus 0
    assert lib.dll.l2b_create_multi(hdr, 9, 1, 1, 0, C.byref(out)) == pkg.capi.EINVAL     # > 8
    assert lib.dll.l2b_create_multi(hdr, 4, 2, 1, 0, C.byref(out)) == pkg.capi.EINVAL     # tp_degree not 1 / n_gpus
    assert b"tp_degree" in lib.dll.l2b_last_error(None)
    assert lib.dll.l2b_create_multi(hdr, 2, 2, 3, 0, C.byref(out)) == pkg.capi.EINVAL     # TP group holds one sequence
    assert lib.dll.l2b_create_multi(None, 1, 1, 1, 0, C.byref(out)) == pkg.capi.EINVAL


def test_no_gpu_means_loud_failure_not_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.L2BError) as e:
        pkg.Context(pkg.synth.header("tiny"))
    assert e.value.code == pkg.capi.ECUDA
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_create_argument_validation_needs_no_gpu(pkg):
    lib = pkg.capi.Library.get()
    out = C.c_void_p()
    bad = (C.c_int32 * 7)(64, 176, 2, 5, 5, 512, 32)         # dim % n_heads != 0
    assert lib.dll.l2b_create(bad, 0, 1, 0, C.byref(out)) == pkg.capi.EINVAL
    assert b"n_heads" in lib.dll.l2b_last_error(None)
    bad = (C.c_int32 * 7)(64, 176, 2, 4, 4, 512, 32)
    assert lib.dll.l2b_create(bad, 0, 0, 0, C.byref(out)) == pkg.capi.EINVAL   # max_batch 0
    assert lib.dll.l2b_create(bad, 0, 1, 33, C.byref(out)) == pkg.capi.EINVAL  # steps > seq_len
    assert lib.dll.l2b_forward(None, 1, 0, None) == pkg.capi.EINVAL
    lib.dll.l2b_destroy(None)                                                # no-op


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under llama2.ts_b200/ may reference it."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama2.ts_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "l2ref" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_public_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/llama2_b200.h must compile as C99 without any C++ / CUDA / torch type."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "h.c"
    src.write_text('#include "include/llama2_b200.h"\nint main(void) { return L2B_ABI_VERSION == 2 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-I", root, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
