"""boost_posterior (clean_phantoms, run_time_info.f90:820-877): the phantoms phase U removes
are promoted to posterior samples with probability boost_posterior / num_repeats, each carrying the weight of the death
since the last update with the smallest logL above its own.  The engine's list against the oracle's on identical seeds,
and the posterior files that hold them."""
import ctypes as C

import numpy as np
import pytest

from polychordlite_b200 import pypolychord

pytestmark = pytest.mark.gpu


def _canon(rows, idx, lw):
    key = np.lexsort((rows[:, -1], idx))   # by (dead index, logL)
    return rows[key], idx[key], lw[key]


def _both(gpu, oracle, like="gaussian", **kw):
    K = kw.pop("batch_K")
    npars = kw["nDims"] + kw.get("nDerived", 0) + 2
    gpu.set_option("batch_K", K)
    try:
        gi, gd = gpu.run(gpu.make_settings(**kw), like=like, want_dump=True)
    finally:
        gpu.set_option("batch_K", 0)
    g = gpu.last_boosted(npars)
    oi, od = oracle.run(oracle.make_settings(batch_K=K, **kw), like=like, want_dump=True)
    o = oracle.last_boosted(npars)
    return gi, gd, g, oi, od, o


@pytest.mark.parametrize("kw", [
    dict(nDims=4, nDerived=1, nlive=64, num_repeats=8, seed=0, batch_K=16, boost_posterior=-1.0, posteriors=True),
    dict(nDims=6, nDerived=0, nlive=100, num_repeats=12, seed=3, batch_K=25, boost_posterior=-1.0, posteriors=True),
    dict(nDims=20, nDerived=2, nlive=200, num_repeats=40, seed=1, batch_K=50, boost_posterior=-1.0, equals=True),
])
def test_every_phantom_promoted_matches_the_oracle(gpu, oracle, kw):
    """boost_posterior < 0: thin_posterior = 1, every removed phantom becomes a sample -- the same samples, taking their
    weights from the same dead points, as the oracle's clean_phantoms."""
    gi, gd, (gr, gx, gw), oi, od, (orows, ox, ow) = _both(gpu, oracle, **dict(kw))
    # the run itself is unchanged by the boost
    assert (gi.ndead, gi.nlike, gi.nupdates, gi.nphantoms_final) == (oi.ndead, oi.nlike, oi.nupdates, oi.nphantoms_final)
    assert abs(gi.logZ - oi.logZ) < 1e-7
    assert len(gx) == len(ox) and len(gx) > 0
    gr, gx, gw = _canon(gr, gx, gw)
    orows, ox, ow = _canon(orows, ox, ow)
    assert np.array_equal(gx, ox)
    # same tolerance as the dead points of the run parity tests (FP re-association in the slice arithmetic)
    assert np.allclose(gr, orows, rtol=0, atol=1e-6)
    assert np.allclose(gw, ow, rtol=0, atol=1e-6)
    # every sample lies below the death it takes its weight from; all phantoms but those still above the last
    # contour when the run ends are promoted
    dl = gd[-1]["dead"][:, -1]
    assert np.all(gr[:, -1] < dl[gx])
    nph = gi.nslices - gi.nchains   # babies that were not the last of their chain
    assert 0.5 * nph < len(gx) <= nph


@pytest.mark.parametrize("kw", [
    dict(nDims=4, nDerived=1, nlive=64, num_repeats=8, seed=0, batch_K=16, boost_posterior=4.0, posteriors=True),
    dict(nDims=20, nDerived=2, nlive=1000, num_repeats=40, seed=2, batch_K=250, boost_posterior=5.0, posteriors=True),  # BASELINE config 2
])
def test_thinned_samples_are_a_subset_of_the_oracles_full_set(gpu, oracle, kw):
    """thin_posterior = boost_posterior / num_repeats < 1.  The engine addresses the Bernoulli trial by the bits of the
    sample's own logL, which agree with the oracle's only to rounding, so the two thin differently; what must agree is
    everything else: each promoted sample is one of the oracle's removed phantoms (its boost_posterior < 0 list) with
    the same dead point and weight, the promoted fraction is thin_posterior, and a repeated run promotes the same set."""
    full = dict(kw, boost_posterior=-1.0)
    gi, gd, (gr, gx, gw), oi, od, (orows, ox, ow) = _both(gpu, oracle, **dict(kw))
    oracle.run(oracle.make_settings(**{**full, "batch_K": kw["batch_K"]}), like="gaussian")
    npars = kw["nDims"] + kw["nDerived"] + 2
    frows, fx, fw = _canon(*oracle.last_boosted(npars))
    gr, gx, gw = _canon(gr, gx, gw)
    # position of every engine sample in the oracle's full list: same dead point, nearest logL
    start = np.searchsorted(fx, gx, side="left")
    stop = np.searchsorted(fx, gx, side="right")
    assert np.all(stop > start)
    pos = np.empty(len(gx), dtype=np.int64)
    for i in range(len(gx)):
        seg = frows[start[i]:stop[i], -1]
        pos[i] = start[i] + int(np.argmin(np.abs(seg - gr[i, -1])))
    assert len(np.unique(pos)) == len(pos)
    assert np.allclose(gr, frows[pos], rtol=0, atol=1e-6)
    assert np.allclose(gw, fw[pos], rtol=0, atol=1e-6)
    thin = kw["boost_posterior"] / kw["num_repeats"]
    nfull = len(fx)
    assert abs(len(gx) - thin * nfull) < 5 * np.sqrt(thin * (1 - thin) * nfull)
    # reproducible
    gpu.set_option("batch_K", kw["batch_K"])
    try:
        k2 = {k: v for k, v in kw.items() if k != "batch_K"}
        gpu.run(gpu.make_settings(**k2))
    finally:
        gpu.set_option("batch_K", 0)
    r2, x2, w2 = _canon(*gpu.last_boosted(npars))
    assert np.array_equal(x2, gx) and np.array_equal(r2, gr) and np.array_equal(w2, gw)


def test_no_boost_without_posterior_files_or_with_zero_boost(gpu):
    for kw in (dict(boost_posterior=4.0), dict(boost_posterior=0.0, posteriors=True)):
        gpu.run(gpu.make_settings(nDims=4, nDerived=1, nlive=64, num_repeats=8, seed=0, **kw))
        assert len(gpu.last_boosted(7)[1]) == 0


def test_boosted_list_grows_past_its_first_allocation(gpu, oracle):
    kw = dict(nDims=4, nDerived=0, nlive=64, num_repeats=8, seed=5, batch_K=16, boost_posterior=-1.0, posteriors=True)
    gpu.set_option("cap_ph0", 16 * 7)   # smallest pools: the boost list (2 x pool) must be regrown several times
    try:
        gi, gd, (gr, gx, gw), oi, od, (orows, ox, ow) = _both(gpu, oracle, **kw)
    finally:
        gpu.set_option("cap_ph0", 0)
    assert len(gx) == len(ox) > 2 * 16 * 7
    assert np.array_equal(np.sort(gx), np.sort(ox))


def test_posterior_files_hold_the_boosted_samples(gpu, tmp_path):
    D, P, n, R = 4, 1, 100, 8
    L = gpu.lib()
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = pypolychord.polychord._ARGTYPES
    gf, gd, comm = (C.c_double * 1)(1.0), (C.c_int * 1)(D), C.c_int(0)

    def call(boost, root):
        L.polychord_c_interface(C.cast(L.pc_gaussian_loglikelihood, C.c_void_p), C.cast(L.pc_unit_prior, C.c_void_p), None,
                                n, R, -1, -1, False, 0, 1e-3, -1e30, -1, boost, True, True, False, False, False, False,
                                False, False, False, False, False, float(np.exp(-1)), True, D, P, str(tmp_path).encode(),
                                root, 1, gf, gd, 0, None, None, 7, C.byref(comm))
        info = gpu.last_run_info()
        assert info.status == 0
        return info, np.loadtxt(tmp_path / (root.decode() + ".txt")), np.loadtxt(tmp_path / (root.decode() + "_equal_weights.txt"))

    i0, post0, eq0 = call(0.0, b"plain")
    i1, post1, eq1 = call(4.0, b"boost")
    nb = len(gpu.last_boosted(D + P + 2)[1])
    assert i0.ndead == i1.ndead and abs(i0.logZ - i1.logZ) < 1e-12
    assert nb > i1.ndead          # about 7/2 promoted phantoms per death
    assert len(post1) > len(post0) + nb // 2 and len(eq1) > len(eq0)
    assert post1[:, 0].max() == 1.0 and post1.shape[1] == 2 + D + P
    # the weighted posterior moments agree between the two files (Gaussian, mu = 0.5, sigma = 0.1)
    for post in (post0, post1):
        w = post[:, 0] / post[:, 0].sum()
        mean = (w[:, None] * post[:, 2:2 + D]).sum(0)
        sd = np.sqrt((w[:, None] * (post[:, 2:2 + D] - mean) ** 2).sum(0))
        assert np.all(np.abs(mean - 0.5) < 0.03) and np.all(np.abs(sd - 0.1) < 0.03)


def test_interrupted_boosted_run_resumes_with_the_same_samples(gpu, tmp_path):
    """The resume file carries the promoted phantoms: an interrupted run picked up from its file
    ends with the list of the uninterrupted one, bit for bit."""
    D, P = 6, 1
    st = gpu.make_settings(D, P, nlive=200, num_repeats=12, seed=11, posteriors=True, boost_posterior=3.0)
    ref, _ = gpu.run(st, want_dump=True)
    r0, x0, w0 = gpu.last_boosted(D + P + 2)
    path = tmp_path / "b.resume"
    gpu.set_option("errors_return", 1)
    try:
        gpu.set_resume(path, write=True)
        gpu.set_option("resume_interval", 0.0)
        with pytest.raises(RuntimeError):
            gpu.run(st, abort_after_dumps=5)
        assert path.exists()
        gpu.set_resume(path, read=True)
        res, _ = gpu.run(st, want_dump=True)
    finally:
        gpu.set_resume()
        gpu.set_option("resume_interval", 1.0)
        gpu.set_option("errors_return", 0)
    r1, x1, w1 = gpu.last_boosted(D + P + 2)
    assert (res.ndead, res.nlike, res.logZ) == (ref.ndead, ref.nlike, ref.logZ)
    assert len(x0) > ref.ndead // 2
    assert np.array_equal(x0, x1) and np.array_equal(r0, r1) and np.array_equal(w0, w1)


def test_host_callback_run_promotes_the_oracles_phantoms(gpu, oracle, tmp_path):
    """Generic host callbacks (row f2): the phantoms carry the derived parameters the user's likelihood returned."""
    D, P, n, R, K = 4, 1, 120, 8, 30
    sig = 0.1

    def callbacks(capi):
        def ll(theta_p, nd, phi_p, nder):
            th = np.ctypeslib.as_array(theta_p, shape=(nd,))
            r2 = float(np.sum(th ** 2))
            phi_p[0] = np.sqrt(r2)
            return -D * (np.log(sig) + 0.5 * np.log(2 * np.pi)) - 0.5 * r2 / sig ** 2

        def prior(cube_p, theta_p, nd):
            for i in range(nd):
                theta_p[i] = -1.0 + 2.0 * cube_p[i]
        return capi.LL_CB(ll), capi.PRIOR_CB(prior)

    ll, prior = callbacks(gpu)
    L = gpu.lib()
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = pypolychord.polychord._ARGTYPES
    gf, gd, comm = (C.c_double * 1)(1.0), (C.c_int * 1)(D), C.c_int(0)
    gpu.set_option("batch_K", K)
    try:
        L.polychord_c_interface(C.cast(ll, C.c_void_p), C.cast(prior, C.c_void_p), None, n, R, -1, -1, False, 0, 1e-3,
                                -1e30, -1, -1.0, True, False, False, False, False, False, False, False, False, False, False,
                                float(np.exp(-1)), True, D, P, str(tmp_path).encode(), b"hc", 1, gf, gd, 0, None, None, 11,
                                C.byref(comm))
    finally:
        gpu.set_option("batch_K", 0)
    info = gpu.last_run_info()
    assert info.status == 0
    gr, gx, gw = _canon(*gpu.last_boosted(D + P + 2))
    oll, oprior = callbacks(oracle)
    oi, _ = oracle.run(oracle.make_settings(D, P, nlive=n, num_repeats=R, seed=11, batch_K=K, posteriors=True,
                                            boost_posterior=-1.0), like="callback", ll_cb=oll, prior_cb=oprior)
    orows, ox, ow = _canon(*oracle.last_boosted(D + P + 2))
    assert (info.ndead, info.nlike) == (oi.ndead, oi.nlike)
    assert len(gx) == len(ox) > 0 and np.array_equal(gx, ox)
    assert np.allclose(gr, orows, rtol=0, atol=1e-6) and np.allclose(gw, ow, rtol=0, atol=1e-6)
    assert np.all(gr[:, D] > 0)          # phi = |theta| came from the callback
