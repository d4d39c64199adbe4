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


def _deps(src):
    """What an object file depends on: its source and, conservatively, every header of the engine."""
    heads = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG.parent / "include").glob("*"))
    if src.suffix == ".cpp":   # host-only units do not see the kernel headers
        heads = [h for h in heads if h.suffix != ".cuh" or h.name == "pc_device.cuh"]
    return [src] + heads


def build(force=False, verbose=False):
    """One object per translation unit (only the stale ones are recompiled, eight at a time), then one link.  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo for every unit."""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    LIBDIR.mkdir(exist_ok=True)
    # the objects are scratch: they live under gpurun_out/ (git-ignored, never shipped to the GPU box -- only the linked
    # library travels), or wherever PC_BUILD_DIR says
    objdir = Path(os.environ.get("PC_BUILD_DIR", str(PKG.parent / "gpurun_out" / ".objs")))
    objdir.mkdir(parents=True, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f not in ("-shared", "--threads", "8")]
    srcs = [p for p in sources() if p.name != SHIM_SRC.name]
    jobs = []
    for src in srcs:
        obj = objdir / (src.stem + ".o")
        if force or not obj.exists() or any(d.stat().st_mtime > obj.stat().st_mtime for d in _deps(src)):
            cmd = [nvcc, *flags, "-I", str(PKG.parent / "include"), "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)
    if verbose:
        print(f"{len(jobs)} of {len(srcs)} units to compile")
    with ThreadPoolExecutor(max_workers=int(os.environ.get("PC_BUILD_JOBS", "8"))) as ex:
        for r in ex.map(lambda c: subprocess.run(c, check=False), jobs):
            if r.returncode != 0:
                raise subprocess.CalledProcessError(r.returncode, r.args)
    objs = [str(objdir / (p.stem + ".o")) for p in srcs]
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB), *objs],
                   check=True)
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
