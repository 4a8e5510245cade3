"""nvcc recipe for libllama2_b200.so (sm_100a only, in-tree output)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "l2b.cu")
DEPS = [os.path.join(HERE, "..", "include", "llama2_b200.h")]   # + every file under csrc/ (see _deps)
OUT = os.path.join(HERE, "libllama2_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--cudart", "static",
]


def _deps():
    d = list(DEPS)
    csrc = os.path.join(HERE, "csrc")
    for f in os.listdir(csrc):
        p = os.path.join(csrc, f)
        if p not in d and f.endswith((".cu", ".cuh", ".h", ".inl")):
            d.append(p)
    return d


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_library(force=False, verbose=False):
    """Compile csrc/l2b.cu -> libllama2_b200.so.  Cross-compiles without a GPU."""
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    # L2B_EXPERIMENTS=1 also compiles the persistent-kernel experiments under experiments/
    # (option "mega"); they are measured slower than the default path and not part of the product
    extra = ["-DL2B_EXPERIMENTS"] if os.environ.get("L2B_EXPERIMENTS") == "1" else []
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
