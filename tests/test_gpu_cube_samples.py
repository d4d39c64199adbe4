"""cube_samples (pypolychord.run, polychord.py:452, 576-579, 650-789): a run that starts from the caller's live points
instead of prior draws -- pc_set_initial_live + polychord_c_interface against the oracle started from the same points,
for a device likelihood and for host callbacks, and the pypolychord keyword."""
import ctypes as C

import numpy as np
import pytest

from polychordlite_b200 import pypolychord

pytestmark = pytest.mark.gpu

D, P, N, R, K = 4, 1, 96, 8, 24


def _c_interface(gpu, ll, prior, dumper, seed):
    L = gpu.lib()
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = pypolychord.polychord._ARGTYPES
    gf, gd, comm = (C.c_double * 1)(1.0), (C.c_int * 1)(D), C.c_int(0)
    gpu.set_option("batch_K", K)
    gpu.set_option("errors_return", 1)
    try:
        L.polychord_c_interface(C.cast(ll, C.c_void_p), C.cast(prior, C.c_void_p), C.cast(dumper, C.c_void_p), N, R, -1, -1,
                                False, 0, 1e-3, -1e30, -1, 0.0, False, False, False, False, False, False, False, False,
                                False, False, False, float(np.exp(-1)), True, D, P, b".", b"cs", 1, gf, gd, 0, None, None,
                                seed, C.byref(comm))
    finally:
        gpu.set_option("batch_K", 0)
        gpu.set_option("errors_return", 0)
    return gpu.last_run_info()


def _cubes(seed=9):
    # a tighter start than the prior: the box 0.3..0.7 around the Gaussian's mean
    return 0.3 + 0.4 * np.random.default_rng(seed).random((N, D))


def test_device_run_from_given_live_points_matches_the_oracle(gpu, oracle):
    cubes = _cubes()
    final = {}

    def dumper(ndead, nlive, npars, live, dead, lw, logZ, logZerr):
        if nlive == 0:
            final["dead"] = np.ctypeslib.as_array(dead, shape=(ndead, npars)).copy()
    dcb = gpu.DUMPER_CB(dumper)
    L = gpu.lib()
    gpu.set_initial_live(cubes)
    info = _c_interface(gpu, L.pc_gaussian_loglikelihood, L.pc_unit_prior, dcb, seed=5)
    assert info.status == 0
    oracle.set_initial_cubes(cubes)
    oi, od = oracle.run(oracle.make_settings(D, P, nlive=N, num_repeats=R, seed=5, batch_K=K), want_dump=True)
    assert (info.ndead, info.nlike, info.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
    assert abs(info.logZ - oi.logZ) < 1e-7
    assert np.allclose(final["dead"], od[-1]["dead"], rtol=0, atol=1e-6)
    # the run really started from the given points: they are the dead points born from the prior (birth = logzero)
    born = final["dead"][final["dead"][:, -2] <= -1e29]
    assert len(born) == N
    assert np.array_equal(np.sort(born[:, :D], axis=0), np.sort(cubes, axis=0))     # unit prior: theta = cube, exactly
    # one-shot: the next run draws from the prior again
    info2 = _c_interface(gpu, L.pc_gaussian_loglikelihood, L.pc_unit_prior, dcb, seed=5)
    born2 = final["dead"][final["dead"][:, -2] <= -1e29]
    assert info2.status == 0 and not np.array_equal(np.sort(born2[:, :D], axis=0), np.sort(cubes, axis=0))


def test_another_number_of_points_is_absorbed_by_the_live_count(gpu, oracle):
    """polychord.py:650-789 writes whatever points it is given into the resume file and the reference's dynamic-nlive rule
    (run_time_info.f90:766-777) brings the count to nlive; here: the batched form of that rule (phase S1)."""
    cubes = _cubes()[:60]
    L = gpu.lib()
    gpu.set_initial_live(cubes)
    info = _c_interface(gpu, L.pc_gaussian_loglikelihood, L.pc_unit_prior, None, seed=5)
    assert info.status == 0
    oracle.set_initial_cubes(cubes)
    oi, _ = oracle.run(oracle.make_settings(D, P, nlive=N, num_repeats=R, seed=5, batch_K=K))
    assert (info.ndead, info.nlike, info.nupdates, info.nchains) == (oi.ndead, oi.nlike, oi.nupdates, oi.nchains)
    assert abs(info.logZ - oi.logZ) < 1e-7


def test_cube_samples_win_over_a_stale_resume_file(gpu, tmp_path):
    """pypolychord.run defaults to read_resume = write_resume = True: an earlier run's <root>.resume must not shadow the
    caller's starting points (the reference overwrites the resume file with them, polychord.py:576-579)."""
    from polychordlite_b200.pypolychord.builtin import Gaussian
    kw = dict(nDerived=P, nlive=N, num_repeats=R, seed=2, feedback=0, base_dir=str(tmp_path), file_root="rs",
              do_clustering=False, _legacy_output=True)
    pypolychord.run(Gaussian(0.5, 0.1, nDerived=P), D, **kw)          # leaves rs.resume behind
    assert (tmp_path / "rs.resume").exists()
    cubes = _cubes(4)
    out = pypolychord.run(Gaussian(0.5, 0.1, nDerived=P), D, cube_samples=cubes, **kw)
    born = out.theta[out.logL_birth <= -1e29]
    assert len(born) == N and np.array_equal(np.sort(born, axis=0), np.sort(cubes, axis=0))


def test_pypolychord_keyword_with_a_python_likelihood(gpu, tmp_path):
    cubes = _cubes(3)

    def likelihood(theta):
        r2 = float(np.sum((np.asarray(theta) - 0.5) ** 2))
        return -0.5 * r2 / 0.01, [np.sqrt(r2)]

    out = pypolychord.run(likelihood, D, nDerived=P, nlive=N, num_repeats=R, cube_samples=cubes, seed=2, feedback=0,
                          base_dir=str(tmp_path), file_root="cs", do_clustering=False, read_resume=False, write_resume=False,
                          _legacy_output=True)
    born = out.theta[out.logL_birth <= -1e29]
    assert len(born) == N and np.array_equal(np.sort(born, axis=0), np.sort(cubes, axis=0))   # default prior: theta = cube
    with pytest.raises(ValueError):   # points of another dimension
        pypolychord.run(likelihood, D, nDerived=P, nlive=N, num_repeats=R, cube_samples=cubes[:, :D - 1], feedback=0,
                        base_dir=str(tmp_path), file_root="cs2", _legacy_output=True)
