"""GPU parity, building blocks: the CUDA engine's device functions (called through the C ABI
probes of include/polychord_b200.h) against the CPU oracle on the same inputs."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
from test_oracle_numerics import PHILOX_KAT  # noqa: E402


def test_philox_known_answers_on_device(gpu):
    for ctr, key, want in PHILOX_KAT:
        assert gpu.device_philox(ctr, key) == want


def test_uniform_stream_bit_exact(gpu, oracle):
    L = oracle.lib()
    for seed, tag, uid, b in [(0, 1, 0, 0), (12345, 5, 2 ** 40 + 17, 3), (2 ** 31 - 1, 3, 2 ** 63 + 5, 9)]:
        got = gpu.device_uniforms(seed, tag, uid, 0, b, 257)
        want = np.array([L.oracle_uniform(seed, tag, uid, a, b) for a in range(257)])
        assert np.array_equal(got, want)  # integer arithmetic + exact conversion: bit-exact
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())
    for case in g["uniforms"]:
        got = gpu.device_uniforms(case["seed"], case["tag"], case["uid"], 0, case["b"], len(case["values"]))
        assert got.tolist() == case["values"]


def test_inv_normal_cdf(gpu, oracle):
    L = oracle.lib()
    p = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 4001), 10.0 ** -np.arange(1, 300, 7.0), [0.5, 0.075, 0.925]])
    got = gpu.device_inv_normal_cdf(p)
    want = np.array([L.oracle_inv_normal_cdf(float(x)) for x in p])
    # tolerance: same AS241 rational functions, FMA contraction differs between g++ and nvcc
    assert np.allclose(got, want, rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("D,R", [(4, 20), (20, 40), (10, 50), (50, 250), (33, 40), (7, 3), (1, 5), (64, 70), (100, 120), (128, 130)])
def test_directions_match_oracle(gpu, oracle, D, R):
    """chordal_sampling.f90:94-145.  The engine projects with lane-parallel classical Gram-Schmidt,
    the oracle (like random_utils.F90:393-396) with the modified form: same basis up to
    O(cond * eps), tolerance 1e-9."""
    s = oracle.make_settings(D, 0, nlive=10, num_repeats=R, seed=3)
    want = oracle.generate_nhats(s, uid=11)
    got = gpu.device_directions(D, R, 3, 11)
    assert np.allclose(got, want, rtol=0, atol=1e-9)
    assert np.allclose((got ** 2).sum(axis=1), 1.0, atol=1e-13)


@pytest.mark.parametrize("n0,count", [(50, 50), (1000, 250), (1000, 1000), (8000, 3000)])
def test_evidence_scan_matches_sequential_recurrences(gpu, oracle, n0, count):
    """run_time_info.f90:211-296 as block-wide scans vs the oracle's sequential logincexp chain."""
    rng = np.random.default_rng(n0 + count)
    logLs = np.sort(rng.normal(-300, 150, count))
    lz = -1e30
    # start from a non-trivial state: first run a prefix through the oracle
    pre = np.sort(rng.normal(-900, 100, 40))
    st0, _ = oracle.evidence_sequence([lz, lz, 0.0, lz, lz, lz, lz, 0.0], pre, n0, 0)
    want, lw = oracle.evidence_sequence(st0.copy(), logLs, n0, 1)
    got, glw = gpu.device_evidence([st0[0], st0[1], st0[2], st0[3], st0[7]], logLs, n0)
    assert np.allclose(got, [want[0], want[1], want[2], want[3], want[7]], rtol=1e-11, atol=1e-10)
    assert np.allclose(glw, lw, rtol=1e-12, atol=1e-11)
    # and from the initial state (logzero everywhere)
    want, lw = oracle.evidence_sequence([lz, lz, 0.0, lz, lz, lz, lz, 0.0], logLs, n0, 1)
    got, glw = gpu.device_evidence([lz, lz, 0.0, lz, 0.0], logLs, n0)
    assert np.allclose(got, [want[0], want[1], want[2], want[3], want[7]], rtol=1e-11, atol=1e-10)


def test_cholesky_and_fallback(gpu, oracle):
    rng = np.random.default_rng(1)
    for D in (1, 3, 20, 50, 64):
        A = rng.normal(size=(D, 2 * D + 3))
        cov = A @ A.T / A.shape[1]
        L, fb = gpu.device_cholesky(cov)
        assert fb == 0
        assert np.allclose(L, oracle.calc_cholesky(cov), rtol=1e-12, atol=1e-14)
    bad = np.array([[1.0, 2.0], [2.0, 1.0]])
    L, fb = gpu.device_cholesky(bad)
    assert fb == 1 and np.allclose(L, np.sqrt(2.0) * np.eye(2))


@pytest.mark.parametrize("like,D,P", [("gaussian", 20, 2), ("gaussian", 4, 0), ("rastrigin", 10, 0),
                                      ("corr_gaussian", 50, 0), ("gaussian", 40, 1)])
def test_calculate_point(gpu, oracle, like, D, P):
    """calculate.f90:6-50 incl. the out-of-cube rule, prior transform, derived parameters."""
    rng = np.random.default_rng(D)
    s_o = oracle.make_settings(D, P, nlive=10, num_repeats=2 * D)
    s_g = gpu.make_settings(D, P, nlive=10, num_repeats=2 * D)
    cubes = rng.uniform(0, 1, size=(300, D))
    cubes[5, 0] = -1e-9
    cubes[6, D - 1] = 1.0 + 1e-9
    cubes[7] = 0.0
    cubes[8] = 1.0
    kw = {}
    if like == "rastrigin":
        kw = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D)
    if like == "corr_gaussian":
        inv, logdet = oracle.random_inverse_covmat(4, D, float(np.float32(0.1)))
        kw = dict(like_params=np.concatenate([np.full(D, 0.5), inv.ravel(order="F"), [logdet]]))
    want, nw = oracle.calculate_points(s_o, cubes, like=like, **kw)
    got, ng = gpu.calculate_points(s_g, cubes, like=like, **kw)
    assert ng == nw
    assert np.array_equal(got[:, :D], want[:, :D])
    assert np.allclose(got[:, D:2 * D], want[:, D:2 * D], rtol=1e-15, atol=0)
    # logL: tolerance 1e-12 relative (reduction order: warp tree vs serial loop)
    assert np.allclose(got[:, -1], want[:, -1], rtol=1e-12, atol=1e-12)
    assert np.allclose(got[:, 2 * D:2 * D + P], want[:, 2 * D:2 * D + P], rtol=1e-12, atol=1e-12)
    assert got[5, -1] == -1e30 and got[6, -1] == -1e30
