"""In-tree build of libchord.so (the C-ABI library) for sm_100a.

`python -m polychordlite_b200._build` or `build()` from `__graft_entry__`.  The library is
built with nvcc directly (no JIT cache): it lands in polychordlite_b200/lib/ so that it
travels with the source tree.
"""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libchord.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "8",
]


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "polychord_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [str(p) for p in sources() if p.name != "_pypolychord.cpp"]
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(PKG.parent / "include"), "-o", str(LIB), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
