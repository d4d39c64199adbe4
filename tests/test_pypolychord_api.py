"""The host-side mirror of the reference's Python API (polychordlite_b200.pypolychord).

These follow the reference's own tests/test_run_pypolychord.py: the 4-D spherical Gaussian (sigma = 0.1) with
UniformPrior(-1, 1), runs complete, seed >= 0 reproduces bit for bit and seed < 0 does not (:77-90), results
do not depend on nDerived (:93-119), grade_dims must sum to nDims (:122-130) -- plus the numeric check the
reference lacks: logZ against the analytic value -4 ln 2.
"""
import numpy as np
import pytest

from polychordlite_b200 import pypolychord
from polychordlite_b200.pypolychord.builtin import Gaussian, Rastrigin
from polychordlite_b200.pypolychord.priors import (GaussianPrior, LogUniformPrior, SortedUniformPrior, UniformPrior,
                                                    forced_indentifiability_transform)
from polychordlite_b200.pypolychord.settings import PolyChordSettings

nDims = 4
KW = dict(nlive=200, num_repeats=12, feedback=0, do_clustering=False, write_resume=False, read_resume=False,
          write_stats=False, write_live=False, write_dead=False, write_prior=False, posteriors=False, equals=False)


# ----------------------------------------------------------------------------- CPU (no device needed)
def test_unknown_keyword_is_a_type_error(tmp_path):
    with pytest.raises(TypeError):
        pypolychord.run(Gaussian(), nDims, base_dir=str(tmp_path), not_a_keyword=1)


def test_grade_dims_must_sum_to_nDims(tmp_path):
    with pytest.raises(ValueError):
        pypolychord.run(Gaussian(), nDims, base_dir=str(tmp_path), grade_dims=[1, 2])
    with pytest.raises(ValueError):
        PolyChordSettings(nDims, 0, grade_dims=[1, 2])
    with pytest.raises(TypeError):
        PolyChordSettings(nDims, 0, nonsense=True)


def test_run_creates_directories_and_paramnames(tmp_path):
    base = tmp_path / "chains"
    try:  # without a device the engine refuses to run (no CPU fallback) -- after the reference's own preamble has run
        pypolychord.run(lambda theta: -float(np.sum(theta ** 2)), nDims, base_dir=str(base), file_root="t", nlive=20,
                        num_repeats=4, feedback=0, paramnames=[("a", "\\alpha"), ("b", "\\beta")])
    except RuntimeError:
        pass
    assert (base / "clusters").is_dir()
    assert (base / "t.paramnames").read_text().splitlines() == ["a   \\alpha", "b   \\beta"]


def test_settings_defaults_follow_the_reference():
    s = PolyChordSettings(5, 2)
    assert (s.nlive, s.num_repeats, s.do_clustering, s.precision_criterion, s.logzero) == (125, 25, True, 0.001, -1e30)
    assert s.grade_dims == [5] and s.grade_frac == [1.0] and s.seed == -1 and s.cluster_dir.endswith("clusters")
    assert np.isclose(s.compression_factor, np.exp(-1))


def test_priors_match_their_closed_forms():
    x = np.array([0.1, 0.5, 0.9])
    assert np.allclose(UniformPrior(-1, 1)(x), [-0.8, 0.0, 0.8])
    assert np.allclose(LogUniformPrior(1, 100)(x), [10 ** 0.2, 10.0, 10 ** 1.8])
    from scipy.stats import norm
    assert np.allclose(GaussianPrior(1.0, 2.0)(x), norm.ppf(x, 1.0, 2.0))
    t = forced_indentifiability_transform(np.array([0.3, 0.6, 0.9]))
    assert np.all(np.diff(t) > 0) and np.all((t > 0) & (t < 1))
    assert np.all(np.diff(SortedUniformPrior(0, 10)(np.array([0.3, 0.6, 0.9]))) > 0)
    assert np.allclose(UniformPrior(-1, 1).device_params(2), [-1, -1, 1, 1])


def test_builtin_likelihoods_are_callables():
    logL, phi = Gaussian(mu=0.0, sigma=0.1, nDerived=2)(np.zeros(4))
    assert np.isclose(logL, -4 * (np.log(0.1) + 0.5 * np.log(2 * np.pi))) and np.isclose(phi[0], 0.0)
    assert np.isclose(Rastrigin()(np.zeros(3)), -3 * (np.log(4991.21750) - 10.0))


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_run_returns_samples_with_the_analytic_evidence(gpu, tmp_path):
    seen = []

    def dumper(live, dead, logweights, logZ, logZerr):
        seen.append((live.shape, dead.shape, logweights.shape, float(dead[-1, -1])))  # dead[-1] as the reference's test does

    ns = pypolychord.run(Gaussian(mu=0.0, sigma=0.1), nDims, prior=UniformPrior(-1, 1), dumper=dumper, seed=1,
                         base_dir=str(tmp_path), **KW)
    assert isinstance(ns, pypolychord.NestedSamplesLite)
    assert abs(ns.logZ - (-4 * np.log(2.0))) < 5 * max(ns.logZerr, 0.05)
    assert np.all(np.abs(ns.mean()) < 0.03) and np.all(np.abs(ns.std() - 0.1) < 0.03)
    assert np.all(ns.logL >= ns.logL_birth)
    assert seen and seen[-1][0][0] == 0 and seen[-1][1] == (ns.ndead, nDims + 2)   # final call: all dead
    assert all(s[0] == (200, nDims + 2) for s in seen[:-1])                         # updates: nlive live points


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [-1, 0, 1, 2])
def test_seed(gpu, tmp_path, seed):
    a = pypolychord.run(Gaussian(mu=0.0, sigma=0.1), nDims, prior=UniformPrior(-1, 1), seed=seed, base_dir=str(tmp_path), **KW)
    b = pypolychord.run(Gaussian(mu=0.0, sigma=0.1), nDims, prior=UniformPrior(-1, 1), seed=seed, base_dir=str(tmp_path), **KW)
    assert a.equals(b) != (seed < 0)


@pytest.mark.gpu
def test_no_derived(gpu, tmp_path):
    a = pypolychord.run(Gaussian(mu=0.0, sigma=0.1), nDims, prior=UniformPrior(-1, 1), seed=1, base_dir=str(tmp_path), **KW)
    b = pypolychord.run(Gaussian(mu=0.0, sigma=0.1, nDerived=2), nDims, nDerived=2, prior=UniformPrior(-1, 1), seed=1,
                        base_dir=str(tmp_path), **KW)
    assert a.equals(b) and b.phi.shape == (b.ndead, 2)
    assert np.allclose(b.phi[:, 0], np.sqrt((b.theta ** 2).sum(axis=1)))


@pytest.mark.gpu
def test_legacy_run_polychord(gpu, tmp_path):
    settings = PolyChordSettings(nDims, 0, base_dir=str(tmp_path), seed=3, **KW)
    ns = pypolychord.run_polychord(Gaussian(mu=0.0, sigma=0.1), nDims, 0, settings, prior=UniformPrior(-1, 1))
    assert ns.ndead > 1000 and np.isfinite(ns.logZ)


@pytest.mark.gpu
def test_nlives_schedule_and_grades_run(gpu, tmp_path):
    """The `nlives` keyword (polychord.py:452; dynamic nlive, run_time_info.f90:766-777) reaches the engine: above the
    contour -10 the run keeps 400 live points, so the final kill-off adds 400 dead points and more points die in all."""
    KW2 = {**KW, "_legacy_output": True}
    plain = pypolychord.run(Gaussian(mu=0.0, sigma=0.1), nDims, prior=UniformPrior(-1, 1), base_dir=str(tmp_path), seed=1, **KW2)
    dyn = pypolychord.run(Gaussian(mu=0.0, sigma=0.1), nDims, prior=UniformPrior(-1, 1), base_dir=str(tmp_path), seed=1,
                          nlives={-10.0: 400}, **KW2)
    assert dyn.ndead > plain.ndead + 200 and abs(dyn.logZ - plain.logZ) < 0.6
    with pytest.raises(RuntimeError):   # more entries than the engine's schedule holds are reported, not silently dropped
        pypolychord.run(Gaussian(), nDims, base_dir=str(tmp_path), nlives={-float(i): 100 + i for i in range(20)}, **KW)
    ns = pypolychord.run(Gaussian(mu=0.0, sigma=0.1), nDims, prior=UniformPrior(-1, 1), base_dir=str(tmp_path),
                         grade_dims=[1, 3], seed=1, **KW)                                # fast/slow grades run
    assert ns.info["nslices"] == 2 * KW["num_repeats"] * ns.info["nchains"]


@pytest.mark.gpu
def test_output_files_and_PolyChordOutput(gpu, tmp_path):
    """run_polychord() returns a PolyChordOutput parsed from <root>.stats
    (polychord.py:218, output.py:57-99) and the engine writes the files anesthetic/getdist read."""
    s = PolyChordSettings(nDims, 1, nlive=200, num_repeats=12, feedback=0, do_clustering=False, write_resume=False,
                          read_resume=False, base_dir=str(tmp_path), file_root="f1", seed=3)
    seen = {}

    def dumper(live, dead, logweights, logZ, logZerr):
        seen.update(ndead=dead.shape[0], last=dead[-1].copy(), logZ=logZ)

    out = pypolychord.run_polychord(Gaussian(mu=0.0, sigma=0.1, nDerived=1), nDims, 1, s, UniformPrior(-1, 1), dumper)
    assert isinstance(out, pypolychord.PolyChordOutput)
    assert out.ndead == seen["ndead"] and out.nlive == 0 and out.ncluster == 1
    assert abs(out.logZ - seen["logZ"]) < 1e-12 and abs(out.logZ - (-4 * np.log(2))) < 0.5
    assert out.nlike == out.samples.info["nlike"]
    db = out.dead_birth()                                   # theta, phi, logL, birth
    assert db.shape == (out.ndead, nDims + 1 + 2)
    assert np.allclose(db[-1, :nDims + 1], seen["last"][:nDims + 1], rtol=1e-14)
    assert np.all(np.diff(db[:, -2]) >= 0)                  # dead points leave in order of logL
    assert np.all(db[:, -1] <= db[:, -2])                   # born below where they died
    w = out.weighted_posterior()
    assert np.isclose(w[:, 0].max(), 1.0) and out.nposterior == w.shape[0]
    mean = (w[:, 0] / w[:, 0].sum()) @ w[:, 2:2 + nDims]
    assert np.all(np.abs(mean) < 0.03)                      # the posterior is N(0, 0.1^2)
    assert np.allclose(out.means[:nDims], mean, atol=1e-9)
    assert out.equal_weights().shape[0] == out.nequals > 50
    assert (tmp_path / "f1.prior_info").read_text().split()[:3] == ["nprior", "=", "200"]
    assert (tmp_path / "f1_phys_live.txt").exists()
    out.make_paramnames_files([("p%i" % i, "\\theta_%i" % i) for i in range(nDims)] + [("r*", "r")])
    assert len((tmp_path / "f1.paramnames").read_text().splitlines()) == nDims + 1
