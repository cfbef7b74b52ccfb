"""Build libdunet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The library is a plain C-ABI shared object (include/dunet.h); it does not link torch.  It is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdunet_b200.so")
STAMP = os.path.join(HERE, ".libdunet_b200.stamp")
SOURCES = ["dunet.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "dunet.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdunet_b200.so cannot be built (there is no non-CUDA fallback)")


def build(force: bool = False, verbose: bool = False) -> str:
    import fcntl
    import tempfile

    digest = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest

    if not force and fresh():
        return LIB
    with open(os.path.join(HERE, ".libdunet_b200.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)  # one builder at a time (torchrun / mp.spawn ranks start together)
        if not force and fresh():          # another process built it while we waited
            return LIB
        fd, tmp = tempfile.mkstemp(prefix=".libdunet_b200.", suffix=".so.tmp", dir=HERE)
        os.close(fd)
        cmd = [nvcc_path(), *NVCC_FLAGS, "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES], "-lcudart_static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(os.path.join(HERE, "build.log"), "w") as fh:
            fh.write(" ".join(cmd) + "\n" + log)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed building libdunet_b200.so")
        os.replace(tmp, LIB)  # atomic: a concurrent dlopen sees the old or the new file, never a partial one
        if verbose:
            print(log)
        with open(STAMP, "w") as fh:
            fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
