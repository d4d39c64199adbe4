"""The C-ABI library loads without a GPU and exports every symbol include/polychord_b200.h declares."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "polychord_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:void|int|long long|double|char\s*\*|const char\s*\*)\s+\**\s*((?:pc_|polychord_)\w+)\s*\(",
                       text, flags=re.M)
    return sorted(set(names))


def test_header_symbols_are_exported(capi):
    L = capi.lib()
    syms = declared_symbols()
    assert "polychord_c_interface" in syms and "polychord_c_interface_ini" in syms and len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == syms


def test_version_and_options(capi):
    assert b"polychordlite_b200" in capi.lib().pc_version()
    old = capi.get_option("batch_fraction")
    capi.set_option("batch_fraction", 0.5)
    assert capi.get_option("batch_fraction") == 0.5
    capi.set_option("batch_fraction", old)
    try:
        capi.set_option("no_such_option", 1)
        raise AssertionError("unknown option accepted")
    except KeyError:
        pass


def test_host_callbacks_evaluate_reference_formulas(capi):
    """The ready-made host callbacks (same signature as likelihoods/CC/CC_likelihood.cpp:45) are plain
    host code: they must agree with the analytic forms (no GPU involved)."""
    import ctypes as C
    import numpy as np
    L = capi.lib()
    D = 20
    theta = (C.c_double * D)(*([0.4] * D))
    phi = (C.c_double * 2)()
    got = L.pc_gaussian_loglikelihood(theta, D, phi, 2)
    want = -D * (np.log(0.1) + 0.5 * np.log(2 * np.pi)) - 0.5 * D
    assert np.isclose(got, want, rtol=1e-13) and np.isclose(phi[0], np.sqrt(D) * 0.1)
    th = (C.c_double * 3)(0.0, 0.0, 0.0)
    got = L.pc_rastrigin_loglikelihood(th, 3, phi, 0)
    assert np.isclose(got, -3 * (np.log(4991.21750) - 10.0), rtol=1e-13)
    cube = (C.c_double * 3)(0.1, 0.5, 0.9)
    out = (C.c_double * 3)()
    L.pc_unit_prior(cube, out, 3)
    assert list(out) == [0.1, 0.5, 0.9]


def test_no_device_means_loud_failure(capi):
    """Without a CUDA device the compute entry points must fail, not fall back to the CPU."""
    if capi.device_count() > 0:
        return
    capi.set_option("errors_return", 1)
    s = capi.make_settings(4, 0, nlive=20, num_repeats=4)
    try:
        capi.run(s)
        raise AssertionError("pc_run succeeded without a GPU")
    except RuntimeError:
        pass
    finally:
        capi.set_option("errors_return", 0)


def test_reference_facade_links_against_this_library(capi, tmp_path):
    """the reference's own C++ facade (src/polychord/c_interface.cpp) compiled from
    where it lies and linked against this repository's libchord.so with no undefined symbols: the two
    extern "C" entry points are the whole link-time contract.  Needs /root/reference (this container only)."""
    import shutil
    import subprocess
    ref = Path("/root/reference/src/polychord")
    if not (ref / "c_interface.cpp").exists() or shutil.which("g++") is None:
        import pytest
        pytest.skip("reference sources or g++ not available")
    out = tmp_path / "libfacade.so"
    cmd = ["g++", "-std=c++11", "-shared", "-fPIC", "-I", str(ref), str(ref / "c_interface.cpp"),
           "-L", str(capi.LIB_PATH.parent), "-lchord", "-Wl,--no-undefined", "-Wl,-rpath," + str(capi.LIB_PATH.parent),
           "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(out)], capture_output=True, text=True).stdout
    assert "polychord_c_interface" in nm and "polychord_c_interface_ini" in nm


def test_host_only_setters_validate_their_arguments(capi):
    """pc_set_initial_live / pc_last_boosted / pc_maximise are callable without a device and check what they are given."""
    import ctypes as C

    import numpy as np
    L = capi.lib()
    L.pc_set_initial_live.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    a = np.zeros((3, 2))
    assert L.pc_set_initial_live(a.ctypes.data_as(C.POINTER(C.c_double)), 3, 0) == -1
    assert L.pc_set_initial_live(a.ctypes.data_as(C.POINTER(C.c_double)), 3, 2) == 0
    assert L.pc_set_initial_live(None, 0, 0) == 0            # cleared again
    capi.set_initial_live(None)
    try:
        capi.set_initial_live(np.zeros(4))
        raise AssertionError("a 1-d array must be refused")
    except ValueError:
        pass
    rows, idx, lw = capi.last_boosted(5)                     # no run yet in this process (or none that boosted)
    assert len(idx) == len(lw) == rows.shape[0]
    L.pc_maximise.restype = C.c_int
    L.pc_maximise.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int, C.c_int,
                              C.POINTER(C.c_double)]
    assert L.pc_maximise(None, None, 2, 0, -1e30, None, 0, 0, None) == -1
