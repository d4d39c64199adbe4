"""The oracle started from given live points (cube_samples, polychord.py:650-789)."""
import numpy as np


def test_oracle_run_starts_from_the_given_points(oracle):
    D, n = 3, 60
    cubes = 0.35 + 0.3 * np.random.default_rng(1).random((n, D))
    for K in (0, 15):
        oracle.set_initial_cubes(cubes)
        res, dumps = oracle.run(oracle.make_settings(D, 0, nlive=n, num_repeats=6, seed=4, batch_K=K), want_dump=True)
        dead = dumps[-1]["dead"]
        born = dead[dead[:, -2] <= -1e29]
        assert len(born) == n and np.array_equal(np.sort(born[:, :D], axis=0), np.sort(cubes, axis=0))
        # the prior volume the run believes in is still the unit cube: starting inside the bulk overestimates Z
        assert res.logZ > -0.5
        res2, _ = oracle.run(oracle.make_settings(D, 0, nlive=n, num_repeats=6, seed=4, batch_K=K))   # one-shot
        assert res2.nlike != res.nlike
