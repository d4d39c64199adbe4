"""Writes tests/golden/reference_cube_samples.resume with the REFERENCE's own writer of the text resume layout:
/root/reference/pypolychord/polychord.py:650-789 (_make_resume_file -- the file pypolychord hands to the Fortran when a
run starts from the caller's live points), imported from /root/reference (this container only; the fixture travels).

Two of that module's imports do not exist here and are replaced by stand-ins: the compiled `_pypolychord` extension
(not used by _make_resume_file) and the `fortranformat` package, of which the function uses FortranRecordWriter with
exactly three edit descriptors -- (nI12), (nE24.15E3) and (A).  The stand-in formats reals with this repository's
format_e24 (pc_format_e24, checked against the reference's own parser in tests/test_reference_golden.py).

Run:  python tests/golden/make_reference_resume.py
"""
import json
import re
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from polychordlite_b200 import _capi as capi  # noqa: E402


class FortranRecordWriter:
    def __init__(self, fmt):
        m = re.fullmatch(r"\((\d*)(I12|E24\.15E3|A)\)", fmt)
        assert m, fmt
        self.kind = m.group(2)

    def write(self, values):
        if self.kind == "A":
            return "".join(str(v) for v in values)
        if self.kind == "I12":
            return "".join("%12d" % int(v) for v in values)
        return "".join(capi.format_e24(float(v)) for v in values)


sys.modules["fortranformat"] = types.SimpleNamespace(FortranRecordWriter=FortranRecordWriter)
sys.modules["_pypolychord"] = types.ModuleType("_pypolychord")
sys.modules["mpi4py"] = None   # not installed in the reference's single-process use either

from pypolychord.polychord import _make_resume_file  # noqa: E402

D, P, N = 4, 1, 24
MU, SIG = 0.5, 0.1


def loglikelihood(theta):
    r2 = float(np.sum((theta - MU) ** 2))
    return -0.5 * r2 / SIG ** 2 - D * np.log(SIG * np.sqrt(2 * np.pi)), [np.sqrt(r2)]


def prior(cube):
    return np.asarray(cube, dtype=float)


def main():
    out = Path(__file__).parent
    rng = np.random.default_rng(11)
    cubes = rng.random((N, D))
    _make_resume_file(loglikelihood, prior=prior, base_dir=str(out), file_root="reference_cube_samples", cube_samples=cubes,
                      logzero=-1e30, grade_dims=[D], num_repeats=np.array([3 * D]), boost_posterior=0.0)
    (out / "reference_cube_samples.json").write_text(json.dumps({
        "source": "/root/reference/pypolychord/polychord.py:650-789 (_make_resume_file), driven by tests/golden/make_reference_resume.py",
        "nDims": D, "nDerived": P, "nlive": N, "num_repeats": 3 * D, "mu": MU, "sigma": SIG, "cubes": cubes.tolist()}, indent=1))
    print(out / "reference_cube_samples.resume")


if __name__ == "__main__":
    main()
