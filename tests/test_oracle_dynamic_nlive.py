"""Dynamic nlive (replace_point, run_time_info.f90:766-777) in the oracle's reference
schedule: above a contour threshold the number of live points moves to the scheduled target, one point per iteration,
and the evidence stays unbiased.  The engine does not have this row yet: this is its parity
target."""
import numpy as np
import pytest


@pytest.fixture()
def schedule(oracle):
    yield oracle.set_nlives
    oracle.set_nlives(None)


def test_live_count_follows_the_schedule_and_the_evidence_is_unbiased(oracle, schedule):
    D = 4
    schedule({-5.0: 300, 2.0: 50})
    zs, errs = [], []
    for seed in range(8):
        # growth is counted as "failures" by the reference's main loop (nested_sampling.F90:315-319: replace_point
        # returns .false. when nothing was deleted), so nfail must exceed the largest increase
        s = oracle.make_settings(D, 0, nlive=100, num_repeats=8, seed=seed, batch_K=0, nfail=100000)
        res, dumps = oracle.run(s, want_dump=True)
        zs.append(res.logZ); errs.append(res.logZerr)
        for d in dumps[:-1]:
            n, lmin = d["live"].shape[0], d["live"][:, -1].min()
            if lmin <= -5.0:
                assert n == 100
            elif lmin <= 2.0:
                assert 100 <= n <= 300
            else:
                assert 50 <= n <= 300
        counts = [d["live"].shape[0] for d in dumps[:-1]]
        assert max(counts) == 300 and counts[-1] == 50
    # 4-D Gaussian, sigma = 0.1, inside the unit cube: Z = 1 (gaussian.f90:8-9)
    assert abs(np.mean(zs)) < 4 * np.std(zs) / np.sqrt(len(zs)) + 0.05
    assert 0.5 < np.std(zs) / np.mean(errs) < 2.0          # the reported error bar is the scatter


def test_growth_beyond_nfail_ends_the_run_like_the_reference(oracle, schedule):
    schedule({-5.0: 300})
    s = oracle.make_settings(4, 0, nlive=100, num_repeats=8, seed=0, batch_K=0)      # nfail = nlive
    res, _ = oracle.run(s)
    schedule(None)
    full, _ = oracle.run(s)
    assert res.ndead < full.ndead / 2


def test_no_schedule_is_the_plain_run(oracle, schedule):
    s = oracle.make_settings(3, 0, nlive=60, num_repeats=6, seed=2, batch_K=0)
    a, _ = oracle.run(s)
    schedule({1e30: 10})                 # a threshold no contour reaches
    b, _ = oracle.run(s)
    assert (a.ndead, a.nlike, a.logZ) == (b.ndead, b.nlike, b.logZ)


def test_batched_schedule_moves_the_live_count_a_generation_at_a_time(oracle, schedule):
    """The engine's schedule with a moving target: max(n - K, target) after every
    generation; unbiased evidence with an honest error bar; no schedule = the plain batched run."""
    D = 4
    plain, _ = oracle.run(oracle.make_settings(D, 0, nlive=100, num_repeats=8, seed=0, batch_K=25))
    schedule({1e30: 10})
    same, _ = oracle.run(oracle.make_settings(D, 0, nlive=100, num_repeats=8, seed=0, batch_K=25))
    assert (plain.ndead, plain.nlike, plain.logZ) == (same.ndead, same.nlike, same.logZ)
    schedule({-5.0: 300, 2.0: 50})
    zs, errs = [], []
    for seed in range(12):
        res, dumps = oracle.run(oracle.make_settings(D, 0, nlive=100, num_repeats=8, seed=seed, batch_K=25), want_dump=True)
        zs.append(res.logZ); errs.append(res.logZerr)
        counts = [d["live"].shape[0] for d in dumps[:-1]]
        assert counts[0] == 100 and max(counts) == 300 and counts[-1] == 50
        for d in dumps[:-1]:
            if d["live"][:, -1].min() <= -5.0:
                assert d["live"].shape[0] == 100
    assert abs(np.mean(zs)) < 4 * np.std(zs) / np.sqrt(len(zs)) + 0.05
    assert 0.5 < np.std(zs) / np.mean(errs) < 2.0
