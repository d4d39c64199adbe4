"""The oracle's restatement of clean_phantoms' posterior conversion (run_time_info.f90:820-877, boost_posterior):
properties that hold whatever the random stream -- which death a promoted phantom takes its weight from, the weight,
the promoted fraction -- in the reference schedule and in the batched one."""
import numpy as np
import pytest


@pytest.mark.parametrize("batch_K", [0, 40])
def test_promoted_phantoms_take_the_weight_of_the_next_death_above(oracle, batch_K):
    D, P, n, R = 4, 1, 160, 8
    s = oracle.make_settings(D, P, nlive=n, num_repeats=R, posteriors=True, boost_posterior=-1.0, seed=5, batch_K=batch_K)
    res, dumps = oracle.run(s, want_dump=True)
    rows, idx, lw = oracle.last_boosted(D + P + 2)
    dead, deadlw = dumps[-1]["dead"], dumps[-1]["logweights"]
    dl = dead[:, -1]
    assert len(idx) > 0.5 * (res.nslices - res.nchains)          # nearly every phantom is removed before the end
    assert np.all(idx >= 0) and np.all(idx < res.ndead)
    assert np.all(rows[:, -1] < dl[idx])                         # the death lies above the sample ...
    assert np.all(rows[:, -1] > rows[:, -2])                     # ... which lies above its own birth contour
    # weight = log w(death) + logL(sample); the dumper's weights are log w + logL, normalised
    shift = (lw - (deadlw[idx] - dl[idx] + rows[:, -1]))
    assert np.ptp(shift) < 1e-9
    # posterior moments of dead + promoted samples: Gaussian mu = 0.5, sigma = 0.1
    allw = np.concatenate([deadlw, lw - shift[0]])
    allx = np.concatenate([dead[:, :D], rows[:, :D]])
    w = np.exp(allw - allw.max()); w /= w.sum()
    mean = (w[:, None] * allx).sum(0)
    sd = np.sqrt((w[:, None] * (allx - mean) ** 2).sum(0))
    assert np.all(np.abs(mean - 0.5) < 0.02) and np.all(np.abs(sd - 0.1) < 0.02)


def test_thinning_fraction_and_switches(oracle):
    D, n, R = 3, 100, 6
    base = dict(nlive=n, num_repeats=R, seed=2, batch_K=25)
    oracle.run(oracle.make_settings(D, 0, posteriors=True, boost_posterior=-1.0, **base))
    nfull = len(oracle.last_boosted(D + 2)[1])
    r, _ = oracle.run(oracle.make_settings(D, 0, equals=True, boost_posterior=3.0, **base))
    nthin = len(oracle.last_boosted(D + 2)[1])
    thin = 3.0 / R                                               # generate.F90:311-316
    assert abs(nthin - thin * nfull) < 5 * np.sqrt(thin * (1 - thin) * nfull)
    r0, _ = oracle.run(oracle.make_settings(D, 0, boost_posterior=3.0, **base))   # no posterior files asked for
    assert len(oracle.last_boosted(D + 2)[1]) == 0
    oracle.run(oracle.make_settings(D, 0, posteriors=True, boost_posterior=0.0, **base))
    assert len(oracle.last_boosted(D + 2)[1]) == 0
    assert (r0.ndead, r0.nlike, r0.logZ) == (r.ndead, r.nlike, r.logZ)   # the boost does not touch the run
