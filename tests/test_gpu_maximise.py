"""`maximise` through polychord_c_interface (maximiser.F90:31-77): after the run the likelihood and the posterior are
maximised from the final live points and <root>.maximum is written in write_max_file's layout
(read_write.F90:754-807)."""
import ctypes as C

import numpy as np
import pytest

from polychordlite_b200 import pypolychord

pytestmark = pytest.mark.gpu


def _numbers(line):
    return [float(line[i:i + 24]) for i in range(0, len(line.rstrip("\n")), 24)]


def test_maximum_file_of_a_gaussian_run(gpu, tmp_path):
    D, P, n, R = 5, 2, 150, 10
    L = gpu.lib()
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = pypolychord.polychord._ARGTYPES
    gf, gd, comm = (C.c_double * 1)(1.0), (C.c_int * 1)(D), C.c_int(0)
    L.polychord_c_interface(C.cast(L.pc_gaussian_loglikelihood, C.c_void_p), C.cast(L.pc_unit_prior, C.c_void_p), None,
                            n, R, -1, -1, False, 0, 1e-3, -1e30, -1, 0.0, True, False, False, False, False, False,
                            False, False, False, False, True, float(np.exp(-1)), True, D, P, str(tmp_path).encode(),
                            b"mx", 1, gf, gd, 0, None, None, 3, C.byref(comm))
    info = gpu.last_run_info()
    assert info.status == 0
    lines = (tmp_path / "mx.maximum").read_text().split("\n")
    assert lines[0] == "Maximum LogLikelihood:" and lines[2] == "Maximum Likelihood point:"
    assert lines[5] == "Maximum Posterior:" and lines[7] == "Maximum Likelihood at posterior:"
    assert lines[9] == "Maximum Posterior point:" and lines[12] == "LogLikelihood(mean):" and lines[14] == "mean point:"
    peak = -D * (np.log(0.1) + 0.5 * np.log(2 * np.pi))            # gaussian.f90: mu = 0.5, sigma = 0.1
    (maxl,), point = _numbers(lines[1]), _numbers(lines[3])
    assert len(point) == D + P
    assert abs(maxl - peak) < 1e-3 and np.allclose(point[:D], 0.5, atol=3e-3)
    assert point[D] == pytest.approx(np.linalg.norm(np.array(point[:D]) - 0.5), abs=1e-9)   # phi_1 = |theta - mu|
    # uniform prior on the unit cube: log density 0, so the posterior maximum is the likelihood's
    (maxp,), (lpost,), ppoint = _numbers(lines[6]), _numbers(lines[8]), _numbers(lines[10])
    assert abs(maxp - lpost) < 1e-6 and abs(lpost - peak) < 1e-3 and np.allclose(ppoint[:D], 0.5, atol=3e-3)
    (lmean,), mean = _numbers(lines[13]), _numbers(lines[15])
    assert np.allclose(mean[:D], 0.5, atol=0.03) and peak - 0.5 * D * (0.03 / 0.1) ** 2 < lmean <= peak
