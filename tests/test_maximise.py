"""The `maximise` setting, host side (csrc/pc_maximise.cpp; replaces maximiser.F90 + nelder_mead.f90): pc_maximise and
pc_prior_log_density against the numpy restatement in oracle/maximise_oracle.py and against analytic maxima."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
import maximise_oracle as mo  # noqa: E402

from polychordlite_b200 import _capi  # noqa: E402

D, P = 3, 1
MU, SIG = np.array([0.3, 0.55, 0.7]), np.array([0.05, 0.1, 0.2])
LOGZERO = -1e30


def _loglike(theta):
    return float(-0.5 * np.sum(((theta - MU) / SIG) ** 2))


def _prior_sq(cube):                       # theta = cube^2: density 1 / (2 sqrt(theta)), falls with theta
    return np.asarray(cube) ** 2


def _prior_unit(cube):
    return np.asarray(cube) * 1.0


def _c_callbacks(prior_fn):
    def ll(theta_p, nd, phi_p, nder):
        th = np.ctypeslib.as_array(theta_p, shape=(nd,))
        if nder > 0:
            phi_p[0] = float(np.sum(th))
        return _loglike(th)

    def prior(cube_p, theta_p, nd):
        th = prior_fn(np.ctypeslib.as_array(cube_p, shape=(nd,)))
        for i in range(nd):
            theta_p[i] = th[i]
    return _capi.LL_CB(ll), _capi.PRIOR_CB(prior)


def _live(prior_fn, n=40, seed=4):
    rng = np.random.default_rng(seed)
    T = 2 * D + P + 2
    rec = np.zeros((n, T))
    rec[:, :D] = rng.random((n, D))
    for r in rec:
        r[D:2 * D] = prior_fn(r[:D])
        r[2 * D] = np.sum(r[D:2 * D])
        r[T - 1] = _loglike(r[D:2 * D])
    return rec


def _maximise(prior_fn, rec, posterior):
    L = _capi.lib()
    L.pc_maximise.restype = C.c_int
    L.pc_maximise.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int, C.c_int,
                              C.POINTER(C.c_double)]
    ll, prior = _c_callbacks(prior_fn)
    out = np.zeros(rec.shape[1])
    rc = L.pc_maximise(C.cast(ll, C.c_void_p), C.cast(prior, C.c_void_p), D, P, LOGZERO,
                       np.ascontiguousarray(rec).ctypes.data_as(C.POINTER(C.c_double)), len(rec), int(posterior),
                       out.ctypes.data_as(C.POINTER(C.c_double)))
    return rc, out


def test_determinant_free_prior_density():
    L = _capi.lib()
    L.pc_prior_log_density.restype = C.c_double
    L.pc_prior_log_density.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    _, prior = _c_callbacks(_prior_sq)
    for cube in (np.array([0.2, 0.5, 0.9]), np.array([0.999999, 0.3, 0.6])):   # the second steps backwards at the edge
        got = L.pc_prior_log_density(C.cast(prior, C.c_void_p), cube.ctypes.data_as(C.POINTER(C.c_double)), D)
        assert abs(got - mo.dXdtheta(_prior_sq, cube)) < 1e-9
        assert abs(got - (-np.sum(np.log(2 * cube)))) < 1e-3      # dX/dtheta = prod 1 / (2 cube)


@pytest.mark.parametrize("prior_fn,posterior", [(_prior_unit, False), (_prior_sq, False), (_prior_sq, True)])
def test_maximum_matches_the_restatement_and_the_analytic_answer(prior_fn, posterior):
    rec = _live(prior_fn)
    rc, out = _maximise(prior_fn, rec, posterior)
    assert rc == 0
    T = rec.shape[1]
    want = mo.do_maximisation(_loglike, prior_fn, rec[:, :D], rec[:, T - 1], LOGZERO, posterior)
    np.testing.assert_allclose(out[:D], want, rtol=0, atol=1e-9)          # the same simplex walk
    theta = out[D:2 * D]
    np.testing.assert_allclose(theta, prior_fn(out[:D]), rtol=0, atol=1e-14)
    assert out[2 * D] == pytest.approx(np.sum(theta)) and out[T - 1] == pytest.approx(_loglike(theta))
    if not posterior:
        np.testing.assert_allclose(theta, MU, rtol=0, atol=2e-3)          # the likelihood's mode
    else:
        # posterior density in theta: exp(logL) / (2 sqrt(theta)) per dimension -> the mode solves
        # (theta - mu) / sigma^2 + 1 / (2 theta) = 0
        mode = 0.5 * (MU + np.sqrt(MU ** 2 - 2 * SIG ** 2))
        np.testing.assert_allclose(theta, mode, rtol=0, atol=3e-3)


def test_no_simplex():
    rec = _live(_prior_unit, n=D)              # fewer than nDims + 1 live points
    assert _maximise(_prior_unit, rec, False)[0] == 1
    rec = _live(_prior_unit)
    rec[:, -1] = LOGZERO
    assert _maximise(_prior_unit, rec, False)[0] == 1


@pytest.mark.parametrize("dims,seed", [(2, 0), (5, 1), (8, 2)])
def test_simplex_walk_on_random_quadratics(dims, seed):
    """The C++ Nelder-Mead and the numpy restatement make the same moves (reflection / expansion / contraction /
    shrink, the stopping rule with its single-precision exponent) on correlated quadratics in several dimensions."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((dims, dims))
    H = A @ A.T / dims + 0.5 * np.eye(dims)
    centre = rng.uniform(0.3, 0.7, dims)

    def loglike(theta):
        d = np.asarray(theta) - centre
        return float(-0.5 * d @ H @ d / 0.01)

    T = 2 * dims + 2
    n = 4 * dims
    rec = np.zeros((n, T))
    rec[:, :dims] = rng.random((n, dims))
    rec[:, dims:2 * dims] = rec[:, :dims]
    rec[:, T - 1] = [loglike(c) for c in rec[:, :dims]]

    def ll(theta_p, nd, phi_p, nder):
        return loglike(np.ctypeslib.as_array(theta_p, shape=(nd,)))

    def prior(cube_p, theta_p, nd):
        for i in range(nd):
            theta_p[i] = cube_p[i]
    cll, cprior = _capi.LL_CB(ll), _capi.PRIOR_CB(prior)
    L = _capi.lib()
    L.pc_maximise.restype = C.c_int
    L.pc_maximise.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int, C.c_int,
                              C.POINTER(C.c_double)]
    out = np.zeros(T)
    rc = L.pc_maximise(C.cast(cll, C.c_void_p), C.cast(cprior, C.c_void_p), dims, 0, LOGZERO,
                       rec.ctypes.data_as(C.POINTER(C.c_double)), n, 0, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    want = mo.do_maximisation(loglike, _prior_unit, rec[:, :dims], rec[:, T - 1], LOGZERO, False)
    np.testing.assert_allclose(out[:dims], want, rtol=0, atol=1e-9)
    # where the walk stops: the values at the vertices agree to 1e-5, i.e. |x - centre| ~ sqrt(2e-5 * 0.01 / lambda_min)
    assert out[T - 1] > -1e-4 and np.linalg.norm(out[:dims] - centre) < 2e-3
