"""GPU parity, the hot kernel: SliceSampling chains (chordal_sampling.f90:7-92, :163-273) run by
warps on the device against the oracle, chain by chain, on identical (seed point, contour,
Cholesky factor, RNG stream id)."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"

# tolerance on baby coordinates / logL: the chain is a composition of R slice steps; the
# Gram-Schmidt variant (classical, lane-parallel) differs from the oracle's modified form at
# O(cond*eps) and the reductions run in tree order.  Decisions (hence nlike) must be identical.
ATOL = 2e-8


def make_case(oracle, D, P, R, like, nchains, seed, rng, chol_kind="random"):
    s = oracle.make_settings(D, P, nlive=10, num_repeats=R, seed=seed)
    kw = {}
    if like == "rastrigin":
        kw = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D)
        cubes = rng.uniform(0.3, 0.7, size=(nchains, D))
    elif like == "corr_gaussian":
        inv, logdet = oracle.random_inverse_covmat(4, D, float(np.float32(0.1)))
        kw = dict(like_params=np.concatenate([np.full(D, 0.5), inv.ravel(order="F"), [logdet]]))
        cov = np.linalg.inv(inv)
        cubes = 0.5 + rng.multivariate_normal(np.zeros(D), cov, size=nchains)
    else:
        cubes = 0.5 + 0.08 * rng.standard_normal((nchains, D))
    cubes = np.clip(cubes, 1e-6, 1 - 1e-6)
    rec, _ = oracle.calculate_points(s, cubes, like=like, **kw)
    if chol_kind == "identity":
        chol = np.eye(D)
    elif like == "corr_gaussian":
        chol = np.linalg.cholesky(np.linalg.inv(inv))
    else:
        chol = np.tril(rng.standard_normal((D, D)) * 0.01) + 0.05 * np.eye(D)
    logL = rec[:, -1] - rng.uniform(0.5, 5.0, size=nchains)
    uid = np.arange(nchains, dtype=np.uint64) * 7 + 3
    return s, kw, rec, chol, logL, uid


@pytest.mark.parametrize("like,D,P,R,chol_kind", [
    ("gaussian", 20, 2, 40, "random"),      # BASELINE config 2 shape
    ("gaussian", 20, 2, 40, "identity"),    # before the first covariance update (run_time_info.f90:193)
    ("gaussian", 4, 1, 20, "random"),       # the reference's own test problem
    ("rastrigin", 10, 0, 50, "random"),     # config 3 shape
    ("corr_gaussian", 50, 0, 250, "random"),  # config 4 shape (two values per lane, directions in global scratch)
    ("gaussian", 33, 0, 10, "random"),      # ragged: D just above one warp
    ("gaussian", 1, 0, 3, "random"),        # smallest
])
def test_chains_match_oracle(gpu, oracle, like, D, P, R, chol_kind):
    rng = np.random.default_rng(D * 1000 + R)
    nchains = 24 if D >= 50 else 64
    s, kw, rec, chol, logL, uid = make_case(oracle, D, P, R, like, nchains, 17, rng, chol_kind)
    sg = gpu.make_settings(D, P, nlive=10, num_repeats=R, seed=17)
    babies, nlike = gpu.slice_chains(sg, rec, chol, logL, uid, like=like, **kw)
    T = 2 * D + P + 2
    assert babies.shape == (nchains, R, T)
    bad = []
    for c in range(nchains):
        want, nl = oracle.slice_chain(s, rec[c], chol, float(logL[c]), int(uid[c]), like=like, **kw)
        # Rastrigin's gradient (up to 20*pi*10.24 per unit of cube) amplifies the O(eps) differences more
        atol = 1e-6 if like == "rastrigin" else ATOL
        if nl != nlike[c] or not np.allclose(babies[c], want, rtol=0, atol=atol):
            bad.append((c, nl, int(nlike[c]), float(np.abs(babies[c] - want).max())))
        # every baby is inside the contour and carries it as its birth contour
        assert np.all(babies[c][:, -1] >= logL[c])
        assert np.all(babies[c][:, -2] == logL[c])
    assert not bad, bad


def test_chains_with_direction_scratch_in_global_memory(gpu, oracle):
    """Same chains whether the directions are staged in shared memory or in the global scratch."""
    rng = np.random.default_rng(5)
    s, kw, rec, chol, logL, uid = make_case(oracle, 20, 2, 40, "gaussian", 32, 9, rng)
    sg = gpu.make_settings(20, 2, nlive=10, num_repeats=40, seed=9)
    a, na = gpu.slice_chains(sg, rec, chol, logL, uid)
    gpu.set_option("nh_global", 1)
    try:
        b, nb = gpu.slice_chains(sg, rec, chol, logL, uid)
    finally:
        gpu.set_option("nh_global", 0)
    assert np.array_equal(a, b) and np.array_equal(na, nb)


def test_chain_against_golden_fixture(gpu):
    """tests/golden/oracle_golden.json (written by make_golden.py from the oracle)."""
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())
    for case in g["chains"]:
        D, P, R = case["D"], case["P"], case["R"]
        sg = gpu.make_settings(D, P, nlive=10, num_repeats=R, seed=case["seed"])
        kw = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D) if case["like"] == "rastrigin" else {}
        chol = np.array(case["cholesky"]).reshape(D, D)
        babies, nlike = gpu.slice_chains(sg, np.array(case["seed_point"])[None, :], chol, case["logL"],
                                         np.array([case["uid"]], dtype=np.uint64), like=case["like"], **kw)
        assert int(nlike[0]) == case["nlike"]
        assert np.allclose(babies[0, -1], case["last_baby"], rtol=0, atol=ATOL)
        assert np.allclose(babies[0, :, -1], case["babies_logL"], rtol=0, atol=ATOL)


def test_walls_out_of_cube_points_are_never_accepted(gpu):
    """calculate.f90:36-39: a wide bracket near the cube boundary must keep every baby inside [0,1]."""
    D, R = 6, 30
    sg = gpu.make_settings(D, 0, nlive=10, num_repeats=R, seed=1)
    rng = np.random.default_rng(0)
    cubes = rng.uniform(0.0, 0.05, size=(64, D))
    rec, _ = gpu.calculate_points(sg, cubes, like_params=[0.0, 1.0])
    babies, nlike = gpu.slice_chains(sg, rec, 5.0 * np.eye(D), rec[:, -1] - 50.0,
                                     np.arange(64, dtype=np.uint64), like_params=[0.0, 1.0])
    assert np.all(babies[:, :, :D] >= 0.0) and np.all(babies[:, :, :D] <= 1.0)
    assert np.all(babies[:, :, -1] > -1e30)


@pytest.mark.parametrize("D,dims,reps", [(4, [1, 3], [20, 20]), (20, [5, 5, 10], [7, 40, 13]), (9, [8, 1], [3, 5])])
def test_chains_with_parameter_grades_match_oracle(gpu, oracle, D, dims, reps):
    """Fast/slow grades (chordal_sampling.f90:94-145): grade g draws its slice directions in the sub-space of the
    dimensions of grades >= g.  Same babies, same nlike as the oracle; slow dimensions do not move in fast steps."""
    R = sum(reps)
    rng = np.random.default_rng(7)
    so = oracle.make_settings(D, 0, nlive=10, num_repeats=R, seed=3)
    sg = gpu.make_settings(D, 0, nlive=10, num_repeats=R, seed=3)
    cubes = np.clip(0.5 + 0.05 * rng.standard_normal((16, D)), 1e-6, 1 - 1e-6)
    chol = np.tril(rng.standard_normal((D, D)) * 0.01) + 0.05 * np.eye(D)
    gpu.set_grades(dims, reps)
    oracle.set_grades(dims, reps)
    try:
        rec, _ = oracle.calculate_points(so, cubes)
        logL = rec[:, -1] - 2.0
        uid = np.arange(16, dtype=np.uint64) + 100
        babies, nlike = gpu.slice_chains(sg, rec, chol, logL, uid)
        for c in range(16):
            want, nl = oracle.slice_chain(so, rec[c], chol, float(logL[c]), int(uid[c]))
            assert nl == nlike[c]
            assert np.allclose(babies[c], want, rtol=0, atol=2e-8)
        # a step of the fastest grade leaves the dimensions of the slower grades untouched
        steps = np.diff(np.vstack([rec[0, :D], babies[0][:, :D]]), axis=0)
        frozen = (np.abs(steps[:, :dims[0]]).max(axis=1) == 0).sum()
        assert frozen >= sum(reps[1:]) - 1
    finally:
        gpu.set_grades()
        oracle.set_grades()
