"""GPU parity of the DENSE chain phase (csrc/pc_dense.cuh: one chain per point group, 32/G chains per warp, the
sequential state machine of slice_sample in lock step) -- the form an ensemble of runs uses to fill the device.
Same random numbers, same arithmetic as the warp-per-chain phase: checked chain by chain against the oracle
(chordal_sampling.f90:7-92, :163-273) and bit for bit against the warp-per-chain phase, then as whole runs."""
import numpy as np
import pytest

from test_gpu_chains import make_case, ATOL

pytestmark = pytest.mark.gpu


@pytest.fixture
def dense(gpu):
    gpu.set_option("dense", 1)
    yield gpu
    gpu.set_option("dense", 0)


@pytest.mark.parametrize("like,D,P,R,chol_kind,nchains", [
    ("gaussian", 20, 2, 40, "random", 64),     # BASELINE config 2 shape (G=4: 8 chains per warp)
    ("gaussian", 20, 2, 40, "identity", 37),   # ragged: the last warp's groups are partly idle
    ("gaussian", 4, 1, 20, "random", 64),      # the reference's own test problem
    ("rastrigin", 10, 0, 50, "random", 64),    # config 3 shape
    ("gaussian", 33, 0, 10, "random", 9),      # G=8: 4 chains per warp
    ("gaussian", 1, 0, 3, "random", 5),        # smallest
    ("gaussian", 70, 0, 6, "random", 3),       # G=16: 2 chains per warp
])
def test_dense_chains_match_oracle_and_warp_per_chain(gpu, dense, oracle, like, D, P, R, chol_kind, nchains):
    rng = np.random.default_rng(D * 1000 + R)
    s, kw, rec, chol, logL, uid = make_case(oracle, D, P, R, like, nchains, 17, rng, chol_kind)
    sg = gpu.make_settings(D, P, nlive=10, num_repeats=R, seed=17)
    babies, nlike = gpu.slice_chains(sg, rec, chol, logL, uid, like=like, **kw)
    gpu.set_option("dense", 0)
    ref_b, ref_n = gpu.slice_chains(sg, rec, chol, logL, uid, like=like, **kw)
    gpu.set_option("dense", 1)
    # the two device forms make the same evaluations with the same arithmetic
    assert np.array_equal(nlike, ref_n)
    assert np.array_equal(babies, ref_b)
    bad = []
    for c in range(nchains):
        want, nl = oracle.slice_chain(s, rec[c], chol, float(logL[c]), int(uid[c]), like=like, **kw)
        atol = 1e-6 if like == "rastrigin" else ATOL
        if nl != nlike[c] or not np.allclose(babies[c], want, rtol=0, atol=atol):
            bad.append((c, nl, int(nlike[c]), float(np.abs(babies[c] - want).max())))
    assert not bad, bad


@pytest.mark.parametrize("kw,like,extra", [
    (dict(nDims=4, nDerived=1, nlive=64, num_repeats=8, seed=0, batch_K=16), "gaussian", {}),
    (dict(nDims=20, nDerived=2, nlive=200, num_repeats=40, seed=1, batch_K=50), "gaussian", {}),
    (dict(nDims=20, nDerived=2, nlive=1000, num_repeats=40, seed=2, batch_K=250), "gaussian", {}),   # BASELINE config 2
    (dict(nDims=6, nDerived=0, nlive=100, num_repeats=12, seed=3, batch_K=99), "gaussian", {}),      # more chains than one pass holds
    (dict(nDims=2, nDerived=0, nlive=200, num_repeats=6, seed=4, batch_K=50, do_clustering=True), "rastrigin",
     dict(prior_lo=[-5.12] * 2, prior_hi=[5.12] * 2)),
])
def test_dense_run_matches_oracle_batched_mode(gpu, dense, oracle, kw, like, extra):
    kw = dict(kw)
    K = kw.pop("batch_K")
    gpu.set_option("batch_K", K)
    try:
        gi, _ = gpu.run(gpu.make_settings(**kw), like=like, **extra)
    finally:
        gpu.set_option("batch_K", 0)
    oi, _ = oracle.run(oracle.make_settings(batch_K=K, **kw), like=like, **extra)
    assert (gi.ndead, gi.nlike, gi.nchains, gi.ngenerations, gi.nupdates, gi.nfailures) == \
           (oi.ndead, oi.nlike, oi.nchains, oi.ngenerations, oi.nupdates, oi.nfailures)
    assert gi.nphantoms_final == oi.nphantoms_final
    assert abs(gi.logZ - oi.logZ) < 1e-7 and abs(gi.logZerr - oi.logZerr) < 1e-7


def test_ensemble_runs_dense_and_equals_single_runs(gpu):
    """pc_run_ensemble takes the dense chain phase by itself; each member equals the same seed run alone through the
    warp-per-chain phase (same batch size): same deaths, same evaluations, logZ to rounding (the covariance partial
    sums are added over a different number of CTAs)."""
    s = gpu.make_settings(20, 2, nlive=400, num_repeats=40, seed=0)
    gpu.set_option("batch_K", 100)
    try:
        infos = gpu.run_ensemble(s, [3, 5, 9, 11, 12])
        for seed, e in zip([3, 5, 9], infos):
            single, _ = gpu.run(gpu.make_settings(20, 2, nlive=400, num_repeats=40, seed=seed))
            assert (e.ndead, e.nlike, e.nupdates) == (single.ndead, single.nlike, single.nupdates)
            assert abs(e.logZ - single.logZ) < 1e-9
    finally:
        gpu.set_option("batch_K", 0)
    assert infos[0].logZ != infos[1].logZ
