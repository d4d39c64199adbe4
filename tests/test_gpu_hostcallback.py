"""Generic host-callback path: likelihood and prior are arbitrary host functions behind
polychord_c_interface; the chains advance in lock step on the device (csrc/pc_hostchain.cuh) and the calling thread
makes one prior + likelihood call per trial point.  Checked against the CPU oracle driven with the SAME callbacks
(batched schedule): identical ndead / nlike, logZ to rounding -- and against the device-resident path, which must
produce the same run when the callback computes the same function."""
import ctypes as C

import numpy as np
import pytest

from polychordlite_b200 import pypolychord
from polychordlite_b200.pypolychord.priors import UniformPrior

pytestmark = pytest.mark.gpu

D, P = 4, 1
SIG = 0.1


def _gauss(theta):
    r2 = float(np.sum(theta ** 2))
    return -D * (np.log(SIG) + 0.5 * np.log(2 * np.pi)) - 0.5 * r2 / SIG ** 2, [np.sqrt(r2)]


def _c_callbacks(capi, counter):
    def ll(theta_p, nd, phi_p, nder):
        th = np.ctypeslib.as_array(theta_p, shape=(nd,))
        logL, phi = _gauss(th)
        for i in range(nder):
            phi_p[i] = phi[i] if i < len(phi) else 0.0
        counter[0] += 1
        return logL

    def prior(cube_p, theta_p, nd):
        for i in range(nd):
            theta_p[i] = -1.0 + 2.0 * cube_p[i]
    return capi.LL_CB(ll), capi.PRIOR_CB(prior)


def _call_c_interface(capi, ll, prior, dumper, nlive, R, seed, max_ndead=-1):
    L = capi.lib()
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = pypolychord.polychord._ARGTYPES
    gf, gd, comm = (C.c_double * 1)(1.0), (C.c_int * 1)(D), C.c_int(0)
    L.polychord_c_interface(C.cast(ll, C.c_void_p), C.cast(prior, C.c_void_p), C.cast(dumper, C.c_void_p), nlive, R, -1, -1,
                            False, 0, 1e-3, -1e30, max_ndead, 0.0, False, False, False, False, False, False, False, False,
                            False, False, False, float(np.exp(-1)), True, D, P, b".", b"hc", 1, gf, gd, 0, None, None,
                            seed, C.byref(comm))
    return capi.last_run_info()


def test_host_callback_run_matches_oracle_with_the_same_callbacks(gpu, oracle):
    n, R, K = 120, 8, 30
    calls = [0]
    ll, prior = _c_callbacks(gpu, calls)
    dumps = []

    def dumper(ndead, nlive, npars, live, dead, lw, logZ, logZerr):
        dumps.append((ndead, nlive, logZ))
    dcb = gpu.DUMPER_CB(dumper)
    gpu.set_option("batch_K", K)
    try:
        info = _call_c_interface(gpu, ll, prior, dcb, n, R, seed=11)
    finally:
        gpu.set_option("batch_K", 0)
    assert info.status == 0
    ocalls = [0]
    oll, oprior = _c_callbacks(oracle, ocalls)
    oi, _ = oracle.run(oracle.make_settings(D, P, nlive=n, num_repeats=R, seed=11, batch_K=K), like="callback", ll_cb=oll,
                       prior_cb=oprior)
    assert (info.ndead, info.nlike, info.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
    assert abs(info.logZ - oi.logZ) < 1e-7
    assert calls[0] >= info.nlike                 # every counted evaluation was a call into the host function
    assert dumps[-1][1] == 0 and dumps[-1][0] == info.ndead and len(dumps) == info.nupdates + 1
    assert abs(info.logZ - (-D * np.log(2))) < 0.6


def test_python_callables_through_the_pypolychord_api(gpu, tmp_path):
    """The reference's own test shape (tests/test_run_pypolychord.py:10-59): a Python likelihood returning
    (logL, phi), UniformPrior(-1, 1) as a Python callable, a dumper indexing dead[-1]."""
    seen = []

    def likelihood(theta):
        return _gauss(np.asarray(theta))

    class PyUniform:                                  # no device form: forces the prior through the callback too
        def __call__(self, cube):
            return -1.0 + 2.0 * np.asarray(cube)

    def dumper(live, dead, logweights, logZ, logZerr):
        seen.append(dead[-1].copy())

    kw = dict(nDerived=P, nlive=100, num_repeats=8, feedback=0, do_clustering=False, write_resume=False, read_resume=False,
              base_dir=str(tmp_path), file_root="py", seed=2, dumper=dumper)
    s0 = pypolychord.run(likelihood, D, prior=PyUniform(), **kw)
    s1 = pypolychord.run(likelihood, D, prior=UniformPrior(-1, 1), **kw)     # device-form prior, host likelihood
    assert s0.ndead == s1.ndead and abs(s0.logZ - s1.logZ) < 1e-9            # same run either way
    assert seen and abs(s0.logZ - (-D * np.log(2))) < 0.7
    assert np.all(np.abs(s0.mean()) < 0.05) and np.all(np.abs(s0.std() - SIG) < 0.03)
    assert np.allclose(s0.phi[:, 0], np.sqrt(np.sum(s0.theta ** 2, axis=1)))   # derived parameters came through
    out = pypolychord.PolyChordOutput(str(tmp_path), "py")                     # and the files were written
    assert out.ndead == s1.ndead and abs(out.logZ - s1.logZ) < 1e-12


def test_exception_in_python_likelihood_propagates(gpu, tmp_path):
    class Boom(Exception):
        pass
    n = [0]

    def likelihood(theta):
        n[0] += 1
        if n[0] > 300:
            raise Boom("stop")
        return _gauss(np.asarray(theta))[0]

    with pytest.raises(Boom):
        pypolychord.run(likelihood, D, nlive=50, num_repeats=6, feedback=0, do_clustering=False, write_resume=False,
                        read_resume=False, base_dir=str(tmp_path), file_root="boom", seed=1)
    assert n[0] < 2000                                # the run stopped promptly instead of sampling to the end
    # the engine is usable afterwards
    s = pypolychord.run(lambda th: _gauss(np.asarray(th))[0], D, prior=UniformPrior(-1, 1), nlive=50, num_repeats=6,
                        feedback=0, do_clustering=False, write_resume=False, read_resume=False, base_dir=str(tmp_path),
                        file_root="ok", seed=1)
    assert s.ndead > 100
