"""The boundary from the reference's side: libchord.so carries the C++ facade itself
(Settings::Settings, the run_polychord overloads, default_prior / default_dumper -- /root/reference/src/polychord/
c_interface.cpp:6-213, interfaces.hpp:8-92) and ships its own `_pypolychord` extension (pypolychord/_pypolychord.cpp:
119-229).  No GPU needed: nothing here computes."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "polychordlite_b200" / "lib"
REF = Path("/root/reference")


def _exports():
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", str(LIBDIR / "libchord.so")], capture_output=True, text=True, check=True)
    return out.stdout


def test_libchord_exports_the_cxx_facade():
    ex = _exports()
    ll = "double (*)(double*, int, double*, int)"
    pr = "void (*)(double*, double*, int)"
    du = "void (*)(int, int, int, double*, double*, double*, double, double)"
    for sym in ["Settings::Settings(int, int)",
                f"run_polychord({ll}, {pr}, {du}, Settings)",
                f"run_polychord({ll}, {du}, Settings)",
                f"run_polychord({ll}, {pr}, Settings)",
                f"run_polychord({ll}, Settings)",
                f"run_polychord({ll}, void (*)(), std::__cxx11::basic_string<char, std::char_traits<char>, std::allocator<char> >)",
                "default_prior(double*, double*, int)",
                "default_dumper(int, int, int, double*, double*, double*, double, double)",
                "polychord_c_interface", "polychord_c_interface_ini"]:
        assert sym in ex, sym


@pytest.mark.skipif(not (REF / "src/drivers/polychord_CC.cpp").exists(), reason="reference tree not present")
def test_reference_cxx_driver_and_shim_link_against_libchord_alone(tmp_path):
    """The reference's src/drivers/polychord_CC.cpp (+ its example likelihood) and pypolychord/_pypolychord.cpp, compiled
    UNCHANGED from where they lie against the reference's own headers, link with -lchord only."""
    exe = tmp_path / "polychord_CC"
    subprocess.run(["g++", "-std=c++11", "-I", str(REF / "src/polychord"), "-I", str(REF / "likelihoods/CC"),
                    str(REF / "src/drivers/polychord_CC.cpp"), str(REF / "likelihoods/CC/CC_likelihood.cpp"),
                    "-L", str(LIBDIR), "-lchord", "-Wl,--no-undefined", f"-Wl,-rpath,{LIBDIR}", "-o", str(exe)], check=True)
    assert exe.exists()
    import numpy
    import sysconfig
    so = tmp_path / "_pypolychord.so"
    subprocess.run(["g++", "-std=c++11", "-fPIC", "-shared", "-I", str(REF / "src/polychord"),
                    "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
                    str(REF / "pypolychord/_pypolychord.cpp"), "-L", str(LIBDIR), "-lchord", "-o", str(so)], check=True)
    und = subprocess.run(["nm", "-D", "--undefined-only", "-C", str(so)], capture_output=True, text=True, check=True).stdout
    ours = [l for l in und.splitlines() if "run_polychord" in l or "Settings::" in l]
    assert len(ours) == 2, und   # ... and both are defined by libchord.so (previous test)


def test_own_pypolychord_extension_validates_like_the_reference():
    """Our `_pypolychord.run`: 34 positional arguments in the reference's order (polychord.py:600-634); the argument
    checks and messages of _pypolychord.cpp:173-204 fire before anything touches the device."""
    from polychordlite_b200.pypolychord import _pypolychord as shim

    def args(**over):
        a = dict(ll=lambda t, p: 0.0, prior=lambda c, t: None, dumper=lambda *a: None, nDims=3, nDerived=0, nlive=10,
                 num_repeats=3, nprior=-1, nfail=-1, do_clustering=False, feedback=0, precision_criterion=1e-3,
                 logzero=-1e30, max_ndead=-1, boost_posterior=0.0, posteriors=False, equals=False, cluster_posteriors=False,
                 write_resume=False, write_paramnames=False, read_resume=False, write_stats=False, write_live=False,
                 write_dead=False, write_prior=False, maximise=False, compression_factor=0.36787944117144233,
                 synchronous=True, base_dir="chains", file_root="t", grade_frac=[1.0], grade_dims=[3], nlives={}, seed=1)
        a.update(over)
        return tuple(a.values())

    with pytest.raises(ValueError, match="grade_dims must sum to nDims"):
        shim.run(*args(grade_dims=[2]))
    with pytest.raises(ValueError, match="same size"):
        shim.run(*args(grade_frac=[1.0, 1.0]))
    with pytest.raises(TypeError, match="list of integers"):
        shim.run(*args(grade_dims=[3.0]))
    with pytest.raises(TypeError, match="list of doubles"):
        shim.run(*args(grade_frac=["a"]))
    with pytest.raises(TypeError, match="dict mapping floats to integers"):
        shim.run(*args(nlives={1: 2}))
    with pytest.raises(TypeError):
        shim.run(*args()[:-1])   # 33 arguments
    with pytest.raises(TypeError, match="callable"):
        shim.run(*args(ll=3))
