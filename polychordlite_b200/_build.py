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
SHIM_SRC = CSRC / "pypolychord_module.cpp"   # the CPython extension: a separate shared object linked against libchord.so

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
    srcs = [str(p) for p in sources() if p.name != SHIM_SRC.name]
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(PKG.parent / "include"), "-o", str(LIB), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


def shim_path():
    import sysconfig
    return PKG / "pypolychord" / ("_pypolychord" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_shim(force=False):
    """The `_pypolychord` extension module (reference: pypolychord/_pypolychord.cpp, built by setup.py:114-123 with
    -lchord): g++ against Python.h + numpy, linked to the in-tree libchord.so with an $ORIGIN rpath."""
    import sysconfig
    import numpy
    out = shim_path()
    if not force and out.exists() and out.stat().st_mtime > max(SHIM_SRC.stat().st_mtime, LIB.stat().st_mtime):
        return out
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", str(SHIM_SRC),
           "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(), "-L", str(LIBDIR), "-lchord",
           "-Wl,-rpath,$ORIGIN/../lib", "-o", str(out)]   # Python's own symbols resolve against the interpreter at import
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
    print(build_shim(force="--force" in sys.argv))
