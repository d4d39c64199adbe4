"""End-to-end checks of the CPU oracle: the known answers the reference's built-in likelihoods
are normalised to (gaussian.f90:8-9 "evidence of 1.0", rastrigin.f90:33) and the behavioural
contracts of the reference's own tests (tests/test_run_pypolychord.py:77-119)."""
import json
from pathlib import Path

import numpy as np
import pytest

ANALYTIC = json.loads((Path(__file__).parent / "golden" / "analytic.json").read_text())


def ensemble(oracle, nseeds, **kw):
    like = kw.pop("like", "gaussian")
    extra = {k: kw.pop(k) for k in ("like_params", "prior_lo", "prior_hi") if k in kw}
    out = []
    for seed in range(nseeds):
        s = oracle.make_settings(seed=seed, **kw)
        r, _ = oracle.run(s, like=like, **extra)
        out.append(r)
    return out


@pytest.mark.parametrize("batch_K", [0, 1, 50, 100])
def test_logZ_gaussian4_box_matches_analytic(oracle, batch_K):
    """The reference's test problem (tests/test_run_pypolychord.py:10-23): 4-D Gaussian sigma=0.1,
    prior U[-1,1]^4 -> logZ = -4 ln 2.  Ensemble mean within 3.5 standard errors for the reference
    schedule (batch_K=0) and for batched generations."""
    D = 4
    rs = ensemble(oracle, 24, nDims=D, nDerived=1, nlive=200, num_repeats=20, batch_K=batch_K,
                  like_params=[0.0, 0.1], prior_lo=[-1.0] * D, prior_hi=[1.0] * D)
    z = np.array([r.logZ for r in rs])
    se = z.std(ddof=1) / np.sqrt(z.size)
    assert abs(z.mean() - ANALYTIC["gaussian4_box_pm1"]["logZ"]) < 3.5 * se + 0.01
    # the run's own error estimate is the right size
    assert 0.5 < np.mean([r.logZerr for r in rs]) / z.std(ddof=1) < 2.0
    # about 4-6 likelihood calls per slice step
    e = np.mean([r.nlike / r.nslices for r in rs])
    assert 3.0 < e < 8.0


def test_logZ_rastrigin2_matches_analytic(oracle):
    D = 2
    rs = ensemble(oracle, 16, nDims=D, nDerived=0, nlive=400, num_repeats=6, batch_K=100, like="rastrigin",
                  prior_lo=[-5.12] * D, prior_hi=[5.12] * D)
    z = np.array([r.logZ for r in rs])
    se = z.std(ddof=1) / np.sqrt(z.size)
    # clustering is not used here; multimodal but slice sampling with nlive=400 still mixes in 2-D
    assert abs(z.mean() - ANALYTIC["rastrigin2_box_5.12"]["logZ"]) < 4 * se + 0.03


def test_seed_determinism_and_seed_sensitivity(oracle):
    """tests/test_run_pypolychord.py:77-90: same seed >= 0 -> identical output."""
    kw = dict(nDims=4, nDerived=1, nlive=100, num_repeats=12, batch_K=25)
    a, da = oracle.run(oracle.make_settings(seed=2, **kw), want_dump=True)
    b, db = oracle.run(oracle.make_settings(seed=2, **kw), want_dump=True)
    c, _ = oracle.run(oracle.make_settings(seed=3, **kw))
    assert a.logZ == b.logZ and a.nlike == b.nlike and a.ndead == b.ndead
    assert np.array_equal(da[-1]["dead"], db[-1]["dead"])
    assert a.logZ != c.logZ


def test_sampler_path_independent_of_nDerived(oracle):
    """tests/test_run_pypolychord.py:93-119."""
    kw = dict(nDims=4, nlive=100, num_repeats=12, batch_K=25, seed=5)
    a, da = oracle.run(oracle.make_settings(nDerived=0, **kw), want_dump=True)
    b, db = oracle.run(oracle.make_settings(nDerived=2, **kw), want_dump=True)
    assert a.logZ == b.logZ and a.nlike == b.nlike
    assert np.array_equal(da[-1]["dead"][:, :4], db[-1]["dead"][:, :4])
    assert np.array_equal(da[-1]["dead"][:, -1], db[-1]["dead"][:, -1])


def test_final_dump_posterior_moments_and_weights(oracle):
    """dump (nested_sampling.F90:546-590): final call has nlive=0, normalised log-weights; the
    weighted posterior of the 4-D Gaussian has mean 0 and sd 0.1 per dimension."""
    D = 4
    s = oracle.make_settings(nDims=D, nDerived=1, nlive=400, num_repeats=20, batch_K=100, seed=11)
    r, dumps = oracle.run(s, like_params=[0.0, 0.1], prior_lo=[-1.0] * D, prior_hi=[1.0] * D, want_dump=True)
    assert len(dumps) == r.nupdates + 1
    last = dumps[-1]
    assert last["live"].shape[0] == 0 and last["dead"].shape == (r.ndead, D + 1 + 2)
    w = np.exp(last["logweights"])
    assert np.isclose(w.sum(), 1.0, rtol=1e-10)
    theta = last["dead"][:, :D]
    mean = (w[:, None] * theta).sum(0)
    sd = np.sqrt((w[:, None] * (theta - mean) ** 2).sum(0))
    assert np.all(np.abs(mean) < 0.02) and np.all(np.abs(sd - 0.1) < 0.015)
    # birth contours are below the death contours, logL ascending in death order
    assert np.all(last["dead"][:, -2] <= last["dead"][:, -1])
    assert np.all(np.diff(last["dead"][:, -1]) >= 0)
    assert last["logZ"] == r.logZ
    # intermediate dumps carry the live points
    assert dumps[0]["live"].shape == (400, D + 3)


def test_max_ndead_and_counts(oracle):
    s = oracle.make_settings(nDims=4, nDerived=0, nlive=50, num_repeats=8, batch_K=10, max_ndead=95, seed=1)
    r, _ = oracle.run(s)
    assert r.ndead == 95 + 50  # max_ndead deaths in the loop, then the kill-off of the live points
    assert r.nchains == 95 and r.nslices == 95 * 8
    s0 = oracle.make_settings(nDims=4, nDerived=0, nlive=50, num_repeats=8, batch_K=10, max_ndead=0, seed=1)
    r0, _ = oracle.run(s0)
    assert r0.ndead == 50 and r0.nchains == 0 and r0.nlike == 50
