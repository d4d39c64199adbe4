"""Pins the CPU oracle's numeric building blocks to published vectors and numpy/scipy.

The reference (Fortran) cannot be compiled in this image and its own tests hold no numeric
golden vectors (tests/test_run_pypolychord.py asserts behaviour only), so the oracle is pinned
to: Random123's Philox4x32-10 known-answer vectors, scipy's normal quantile (AS241 agrees with
it to ~1e-16 relative), numpy's Cholesky / logaddexp, and a brute-force linear-space restatement
of the evidence recurrences of run_time_info.f90:211-296.
"""
import json
from pathlib import Path

import numpy as np
import pytest
from scipy import special, stats

GOLDEN = Path(__file__).parent / "golden"

PHILOX_KAT = [  # Random123 kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_philox_known_answers(oracle):
    for ctr, key, want in PHILOX_KAT:
        assert oracle.philox(ctr, key) == want


def test_uniform_open_interval_and_moments(oracle):
    L = oracle.lib()
    u = np.array([L.oracle_uniform(7, 5, 123, a, 0) for a in range(20000)])
    assert u.min() > 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 4 * (1 / 12 / u.size) ** 0.5
    assert abs(u.var() - 1 / 12) < 0.003
    # different stream coordinates give different numbers
    assert L.oracle_uniform(7, 5, 123, 0, 0) != L.oracle_uniform(7, 5, 124, 0, 0)
    assert L.oracle_uniform(7, 5, 123, 0, 0) != L.oracle_uniform(8, 5, 123, 0, 0)
    assert L.oracle_uniform(7, 5, 123, 0, 0) != L.oracle_uniform(7, 4, 123, 0, 0)


def test_inv_normal_cdf_matches_scipy(oracle):
    L = oracle.lib()
    p = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 2001), 10.0 ** -np.arange(1, 300, 7.0), [0.5, 0.075, 0.925]])
    got = np.array([L.oracle_inv_normal_cdf(float(x)) for x in p])
    want = stats.norm.ppf(p)
    assert np.allclose(got, want, rtol=1e-13, atol=1e-15)


def test_logaddexp_logsumexp(oracle):
    L = oracle.lib()
    import ctypes as C
    rng = np.random.default_rng(0)
    a, b = rng.normal(0, 50, 200), rng.normal(0, 50, 200)
    got = np.array([L.oracle_logaddexp(float(x), float(y)) for x, y in zip(a, b)])
    assert np.allclose(got, np.logaddexp(a, b), rtol=1e-15, atol=1e-15)
    assert L.oracle_logaddexp(-1e30, 3.0) == 3.0  # logzero is absorbed exactly
    v = rng.normal(-100, 30, 1000)
    got = L.oracle_logsumexp(v.ctypes.data_as(C.POINTER(C.c_double)), v.size)
    assert np.isclose(got, special.logsumexp(v), rtol=1e-14)


def test_cholesky_and_fallback(oracle):
    rng = np.random.default_rng(1)
    for D in (1, 3, 20, 50):
        A = rng.normal(size=(D, 2 * D + 3))
        cov = A @ A.T / A.shape[1]
        L = oracle.calc_cholesky(cov)
        assert np.allclose(L, np.linalg.cholesky(cov), rtol=1e-11, atol=1e-13)
    bad = np.array([[1.0, 2.0], [2.0, 1.0]])  # not positive definite -> sqrt(trace) * I (utils.F90:634-638)
    assert np.allclose(oracle.calc_cholesky(bad), np.sqrt(2.0) * np.eye(2))


def brute_force_evidence(logLs, n0, dec):
    """Linear-space restatement of SURVEY.md appendix C for one cluster."""
    from mpmath import mp, mpf, exp, log
    mp.dps = 60
    Z = Z2 = ZX = mpf(0)
    X = XX = mpf(1)
    n = n0
    w = []
    for l in logLs:
        L = exp(mpf(float(l)))
        w.append(float(log(X / (n + 1))))
        Z_new = Z + X * L / (n + 1)
        Z2 = Z2 + 2 * ZX * L / (n + 1) + 2 * XX * L * L / ((n + 1) * (n + 2))
        ZX = ZX * n / (n + 1) + XX * L * mpf(n) / ((n + 1) * (n + 2))
        X = X * n / (n + 1)
        XX = XX * n / (n + 2)
        Z = Z_new
        n -= dec
    return [float(log(Z)), float(log(Z2)), float(log(X)), float(log(ZX)), float(log(XX))], w


@pytest.mark.parametrize("n0,dec,count", [(50, 0, 300), (50, 1, 50), (1000, 1, 250)])
def test_evidence_recurrences_against_linear_space(oracle, n0, dec, count):
    pytest.importorskip("mpmath")
    rng = np.random.default_rng(n0 + dec)
    logLs = np.sort(rng.normal(-30, 15, count))
    lz = -1e30
    st, lw = oracle.evidence_sequence([lz, lz, 0.0, lz, lz, lz, lz, 0.0], logLs, n0, dec)
    want, w = brute_force_evidence(logLs, n0, dec)
    got = [st[0], st[1], st[2], st[3], st[7]]
    assert np.allclose(got, want, rtol=1e-11, atol=1e-11)
    assert np.allclose(lw, w, rtol=1e-12, atol=1e-12)
    # single cluster: the per-cluster copies track the global ones exactly
    assert st[4] == st[0] and st[5] == st[1] and st[6] == st[3]


def test_evidence_mean_matches_monte_carlo_over_shrinkage(oracle):
    """<Z> from the recurrences equals the Monte-Carlo mean over t_i ~ Beta(n,1) shrinkage factors."""
    rng = np.random.default_rng(5)
    n, count = 20, 60
    logLs = np.sort(rng.normal(0, 2, count))
    lz = -1e30
    st, _ = oracle.evidence_sequence([lz, lz, 0.0, lz, lz, lz, lz, 0.0], logLs, n, 0)
    t = rng.beta(n, 1, size=(200000, count))
    X = np.cumprod(t, axis=1)
    Xprev = np.concatenate([np.ones((t.shape[0], 1)), X[:, :-1]], axis=1)
    Zs = ((Xprev - X) * np.exp(logLs)).sum(axis=1)
    assert np.isclose(np.exp(st[0]), Zs.mean(), rtol=5 * Zs.std() / Zs.mean() / np.sqrt(t.shape[0]) + 1e-3)
    assert np.isclose(np.exp(st[1]), (Zs ** 2).mean(), rtol=0.02)


@pytest.mark.parametrize("D,R", [(4, 20), (20, 40), (10, 50), (50, 250), (7, 3)])
def test_directions_are_orthonormal_bases_in_shuffled_order(oracle, D, R):
    s = oracle.make_settings(D, 0, nlive=10, num_repeats=R, seed=3)
    nh = oracle.generate_nhats(s, uid=11)
    assert nh.shape == (R, D)
    assert np.allclose((nh ** 2).sum(axis=1), 1.0, atol=1e-12)
    # the R columns are ceil(R/D) orthonormal bases (the last truncated), shuffled except column 0
    G = nh @ nh.T
    nb = -(-R // D)
    close_to_01 = np.isclose(np.abs(G), 0, atol=1e-10) | np.isclose(np.abs(G), 1, atol=1e-10)
    if nb == 1:
        assert np.allclose(G, np.eye(R), atol=1e-10)
    else:
        # every direction is orthogonal to at least min(D,R)-1 others
        assert (np.isclose(G, 0, atol=1e-10).sum(axis=1) >= min(D, R - D * (nb - 1)) - 1).all()
    nh2 = oracle.generate_nhats(s, uid=12)
    assert not np.allclose(nh, nh2)


def test_calculate_point_in_cube_rule_and_likelihood_values(oracle):
    """calculate.f90:36-44: out-of-cube points get logzero without a likelihood call and are not counted."""
    D = 20
    s = oracle.make_settings(D, 2, nlive=10, num_repeats=40)
    cubes = np.full((4, D), 0.5)
    cubes[1, 3] = 1.0 + 1e-12
    cubes[2, 0] = -1e-300
    cubes[3] = 0.4
    rec, n = oracle.calculate_points(s, cubes)
    assert n == 2
    T = 2 * D + 4
    assert rec.shape == (4, T)
    want0 = -D * (np.log(0.1) + 0.5 * np.log(2 * np.pi))          # gaussian.f90 at the peak
    assert np.isclose(rec[0, -1], want0, rtol=1e-14)
    assert rec[1, -1] == -1e30 and rec[2, -1] == -1e30
    assert np.isclose(rec[3, -1], want0 - 0.5 * D * 1.0, rtol=1e-13)
    assert np.isclose(rec[3, 2 * D], np.sqrt(D) * 0.1)             # derived: radius
    vn = np.pi ** (D / 2) / special.gamma(1 + D / 2)
    assert np.isclose(rec[3, 2 * D + 1], np.log((np.sqrt(D) * 0.1) ** D * vn))
    # rastrigin at the origin of a [-5.12, 5.12] box (rastrigin.f90:33)
    s2 = oracle.make_settings(10, 0, nlive=10, num_repeats=50)
    rec2, _ = oracle.calculate_points(s2, np.full((1, 10), 0.5), like="rastrigin", prior_lo=[-5.12] * 10,
                                      prior_hi=[5.12] * 10)
    assert np.isclose(rec2[0, -1], -10 * (np.log(4991.21750) - 10.0), rtol=1e-14)
    assert np.allclose(rec2[0, 10:20], 0.0)


def test_random_inverse_covmat_spec(oracle):
    """random_utils.F90:581-614: eigenvalues sigma_j = sigma * 0.01^{(j-1)/(D-1)}, Haar basis."""
    D, sigma = 50, float(np.float32(0.1))
    inv, logdet = oracle.random_inverse_covmat(4, D, sigma)
    assert np.allclose(inv, inv.T, rtol=1e-10)
    ev = np.sort(1 / np.sqrt(np.linalg.eigvalsh(inv)))[::-1]
    want = sigma * 0.01 ** (np.arange(D) / (D - 1))
    assert np.allclose(ev, want, rtol=1e-8)
    assert np.isclose(logdet, 2 * np.log(want).sum(), rtol=1e-12)


def test_golden_fixture_is_reproduced(oracle):
    """tests/golden/oracle_golden.json was written by tests/golden/make_golden.py in the build
    container; the oracle must reproduce it wherever the tests run (tolerance: FMA contraction)."""
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())
    for case in g["uniforms"]:
        L = oracle.lib()
        got = [L.oracle_uniform(case["seed"], case["tag"], case["uid"], a, case["b"]) for a in range(len(case["values"]))]
        assert got == case["values"]
    for case in g["chains"]:
        s = oracle.make_settings(case["D"], case["P"], nlive=10, num_repeats=case["R"], seed=case["seed"])
        seed_pt = np.array(case["seed_point"])
        chol = np.array(case["cholesky"]).reshape(case["D"], case["D"])
        kw = {}
        if case["like"] == "rastrigin":
            kw = dict(prior_lo=[-5.12] * case["D"], prior_hi=[5.12] * case["D"])
        babies, nlike = oracle.slice_chain(s, seed_pt, chol, case["logL"], case["uid"], like=case["like"], **kw)
        assert nlike == case["nlike"]
        assert np.allclose(babies[-1], np.array(case["last_baby"]), rtol=1e-9, atol=1e-11)
    for case in g["runs"]:
        s = oracle.make_settings(case["D"], case["P"], nlive=case["nlive"], num_repeats=case["R"], seed=case["seed"],
                                 batch_K=case["batch_K"])
        r, _ = oracle.run(s)
        assert r.ndead == case["ndead"] and r.nlike == case["nlike"]
        assert np.isclose(r.logZ, case["logZ"], rtol=0, atol=1e-8)
