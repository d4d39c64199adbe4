"""Fixtures generated from the REFERENCE's own Python modules (tests/golden/make_reference_golden.py imports
/root/reference/pypolychord/priors.py and output.py in the build container; only the JSON travels):

* the prior classes of the mirror package, and the .ini driver's C++ transforms of the same families, against the
  values the reference's classes return;
* the <root>.stats file: this repository's writer still produces, byte for byte, the text the reference's
  PolyChordOutput parsed when the fixture was made, and the mirror parser reads it to the same fields."""
import ctypes as C
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest

from polychordlite_b200 import _capi, pypolychord
from polychordlite_b200.pypolychord import priors as mirror

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_python.json").read_text())


def _gen():
    spec = importlib.util.spec_from_file_location("make_reference_golden", Path(__file__).parent / "golden" / "make_reference_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_prior_classes_return_the_references_values():
    cubes = np.array(GOLD["priors"]["cubes"])
    for case in GOLD["priors"]["cases"]:
        p = getattr(mirror, case["class"])(*case["args"])
        for c, want in zip(cubes, case["theta"]):
            np.testing.assert_allclose(p(c), want, rtol=1e-13, atol=1e-14)
    for c, want in zip(cubes, GOLD["priors"]["forced_indentifiability_transform"]):
        np.testing.assert_allclose(mirror.forced_indentifiability_transform(c), want, rtol=1e-14)


@pytest.mark.parametrize("cls,ini_name", [("UniformPrior", "uniform"), ("GaussianPrior", "gaussian"),
                                          ("LogUniformPrior", "log_uniform"), ("SortedUniformPrior", "sorted_uniform")])
def test_ini_transforms_return_the_references_values(tmp_path, cls, ini_name):
    """priors.f90's families as the .ini driver computes them (C++, pc_ini_prior_transform) against the reference's
    Python classes of the same distributions.  Tolerance: AS241 against scipy's erfinv for the Gaussian."""
    case = next(c for c in GOLD["priors"]["cases"] if c["class"] == cls)
    cubes = np.array(GOLD["priors"]["cubes"])
    n = cubes.shape[1]
    a, b = case["args"]
    lines = ["nlive = 10", "num_repeats = 2"] + [f"P : p{i} | p{i} | 1 | {ini_name} | 1 | {a!r} {b!r}" for i in range(n)]
    path = tmp_path / "g.ini"
    path.write_text("\n".join(lines) + "\n")
    L = _capi.lib()
    L.pc_ini_prior_transform.restype = C.c_int
    L.pc_ini_prior_transform.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    for c, want in zip(cubes, case["theta"]):
        c = np.ascontiguousarray(c)
        out = np.zeros(n)
        assert L.pc_ini_prior_transform(str(path).encode(), c.ctypes.data_as(C.POINTER(C.c_double)),
                                        out.ctypes.data_as(C.POINTER(C.c_double)), n) == 0
        np.testing.assert_allclose(out, want, rtol=1e-12, atol=1e-13)


def test_stats_file_is_what_the_references_parser_read(tmp_path):
    D, P, dead, logw, live, kw = _gen().stats_inputs()
    _capi.write_files(tmp_path, "gold", D, P, dead, logw, live, flags=("stats", "posteriors", "equals"), **kw)
    assert (tmp_path / "gold.stats").read_text() == GOLD["stats"]["text"]
    out = pypolychord.PolyChordOutput(str(tmp_path), "gold")
    for key, want in GOLD["stats"]["parsed"].items():
        got = getattr(out, key)
        if isinstance(want, list):
            assert list(got) == want, key
        else:
            assert got == want, key


def test_settings_defaults_are_the_references():
    for case in GOLD["settings"]:
        st = pypolychord.PolyChordSettings(case["nDims"], case["nDerived"], **case["kwargs"])
        got = {k: (float(v) if isinstance(v, np.floating) else v) for k, v in vars(st).items()}
        assert got == case["attributes"]
