"""The live count as run state: dynamic nlive schedules (run_time_info.f90:766-777),
nprior > nlive (generate.F90:142-153 + the trim nested_sampling.F90:201-203), and failed births that do not become live
points and count towards nfail (run_time_info.f90:781-785, nested_sampling.F90:315-319, 407-409) -- the engine against
the oracle's batched schedule on the same seeds."""
import ctypes as C

import numpy as np
import pytest

from polychordlite_b200 import pypolychord

pytestmark = pytest.mark.gpu


def both(gpu, oracle, K, schedule=None, want_dump=False, **kw):
    gpu.set_option("batch_K", K)
    gpu.set_nlives(schedule)
    oracle.set_nlives(schedule)
    try:
        gi, gd = gpu.run(gpu.make_settings(**kw), want_dump=want_dump)
        oi, od = oracle.run(oracle.make_settings(batch_K=K, **kw), want_dump=want_dump)
    finally:
        gpu.set_option("batch_K", 0)
        gpu.set_nlives(None)
        oracle.set_nlives(None)
    return gi, gd, oi, od


def same_run(gi, oi):
    assert (gi.ndead, gi.nlike, gi.nchains, gi.ngenerations, gi.nupdates, gi.nfailures) == \
           (oi.ndead, oi.nlike, oi.nchains, oi.ngenerations, oi.nupdates, oi.nfailures)
    assert abs(gi.logZ - oi.logZ) < 1e-7 and abs(gi.logZerr - oi.logZerr) < 1e-7


@pytest.mark.parametrize("schedule", [
    {-40.0: 150, -5.0: 60},      # grows, then shrinks
    {-1e29: 40},                 # shrinks at once: generations of deaths with few births
    {-60.0: 300},                # more than two batches of births in one step
])
def test_dynamic_nlive_schedule_matches_oracle(gpu, oracle, schedule):
    gi, gd, oi, od = both(gpu, oracle, 25, schedule, want_dump=True, nDims=6, nDerived=1, nlive=100, num_repeats=12, seed=3)
    same_run(gi, oi)
    assert [d["live"].shape for d in gd] == [d["live"].shape for d in od]      # the live count follows the schedule ...
    assert len({d["live"].shape[0] for d in gd}) > 1                           # ... and it moved
    for a, b in zip(gd, od):
        assert np.allclose(a["live"], b["live"], rtol=0, atol=1e-6)
        assert np.allclose(a["dead"], b["dead"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("nprior,K", [(300, 25), (1000, 50), (101, 99)])
def test_nprior_points_then_trim_matches_oracle(gpu, oracle, nprior, K):
    gi, _, oi, _ = both(gpu, oracle, K, nDims=5, nDerived=0, nlive=100, num_repeats=10, seed=7, nprior=nprior)
    same_run(gi, oi)
    assert gi.ndead > nprior - 100


def _stepped(capi, step):
    """A likelihood with plateaus: a baby on the contour's own plateau is inside the slice (>=) but does not replace a
    live point (>): a failed birth."""
    def ll(theta_p, nd, phi_p, nder):
        th = np.ctypeslib.as_array(theta_p, shape=(nd,))
        return float(np.floor(-0.5 * np.sum((th - 0.5) ** 2) / 0.01 / step) * step)

    def prior(cube_p, theta_p, nd):
        for i in range(nd):
            theta_p[i] = cube_p[i]
    return capi.LL_CB(ll), capi.PRIOR_CB(prior)


def _c_interface(capi, ll, prior, nlive, R, seed, nfail=-1, max_ndead=-1, D=3):
    L = capi.lib()
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = pypolychord.polychord._ARGTYPES
    gf, gd, comm = (C.c_double * 1)(1.0), (C.c_int * 1)(D), C.c_int(0)
    capi.set_option("errors_return", 1)
    try:
        L.polychord_c_interface(C.cast(ll, C.c_void_p), C.cast(prior, C.c_void_p), None, nlive, R, -1, nfail, False, 0, 1e-3,
                                -1e30, max_ndead, 0.0, False, False, False, False, False, False, False, False, False, False,
                                False, float(np.exp(-1)), True, D, 0, b".", b"dyn", 1, gf, gd, 0, None, None, seed,
                                C.byref(comm))
    finally:
        capi.set_option("errors_return", 0)
    return capi.last_run_info()


@pytest.mark.parametrize("step,nfail", [(0.5, -1), (4.0, 20)])
def test_failed_births_leave_the_live_set_and_count_towards_nfail(gpu, oracle, step, nfail):
    n, R, K = 80, 6, 20
    ll, prior = _stepped(gpu, step)
    gpu.set_option("batch_K", K)
    try:
        info = _c_interface(gpu, ll, prior, n, R, seed=4, nfail=nfail, max_ndead=3000)
    finally:
        gpu.set_option("batch_K", 0)
    assert info.status == 0
    oll, oprior = _stepped(oracle, step)
    oi, _ = oracle.run(oracle.make_settings(3, 0, nlive=n, num_repeats=R, seed=4, batch_K=K, nfail=nfail, max_ndead=3000),
                       like="callback", ll_cb=oll, prior_cb=oprior)
    assert info.nfailures > 0
    assert (info.ndead, info.nlike, info.nchains, info.nfailures) == (oi.ndead, oi.nlike, oi.nchains, oi.nfailures)
    assert abs(info.logZ - oi.logZ) < 1e-7
