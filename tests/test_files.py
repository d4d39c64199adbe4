"""Output files in the reference's formats: the writer is host-only code inside
libchord.so, driven here with synthetic arrays -- no device needed.  Formats: src/polychord/read_write.F90
(write_stats_file :809-910, write_dead_points :679-716, write_phys_live_points :621-677, write_posterior_file
:479-612), numbers in Fortran's E24.15E3 (utils.F90:19)."""
import numpy as np
import pytest

from polychordlite_b200 import _capi as capi
from polychordlite_b200.pypolychord.output import PolyChordOutput


def test_fortran_e24_15e3_edit_descriptor():
    # what gfortran's write(*,'(E24.15E3)') prints for these values
    assert capi.format_e24(1.0) == "  0.100000000000000E+001"
    assert capi.format_e24(-1.0) == " -0.100000000000000E+001"
    assert capi.format_e24(0.0) == "  0.000000000000000E+000"
    assert capi.format_e24(-1e30) == " -0.100000000000000E+031"
    assert capi.format_e24(0.5) == "  0.500000000000000E+000"
    assert capi.format_e24(3.14159265358979e-5) == "  0.314159265358979E-004"
    assert capi.format_e24(123456.789) == "  0.123456789000000E+006"
    # rounding carries into the exponent
    assert capi.format_e24(9.999999999999999e22) == "  0.100000000000000E+024"
    rng = np.random.default_rng(0)
    for v in np.concatenate([rng.standard_normal(200) * 10.0 ** rng.integers(-200, 200, 200), [1e-300, 1e300]]):
        s = capi.format_e24(v)
        assert len(s) == 24 and s[19] == "E"
        assert abs(float(s) - v) <= 5.1e-15 * abs(v)   # 15 significant digits


@pytest.fixture()
def synthetic(tmp_path):
    rng = np.random.default_rng(3)
    D, P, ndead, nlive = 3, 2, 400, 25
    def rows(n, lo):
        theta = rng.uniform(-1, 1, (n, D))
        phi = rng.uniform(0, 1, (n, P))
        logL = np.sort(lo + rng.uniform(0, 5, n))
        birth = np.concatenate([[-1e30] * min(n, 20), logL[:max(n - 20, 0)] - 0.1])
        return np.column_stack([theta, phi, birth, logL])
    dead, live = rows(ndead, -10.0), rows(nlive, -4.0)
    logw = -np.arange(ndead) / 25.0 + dead[:, -1]          # log weight + logL
    nfiles = capi.write_files(tmp_path, "syn", D, P, dead, logw, live, logZ=-3.25, logZerr=0.125, nlike=123456,
                              num_repeats=15, seed=7)
    assert nfiles == 7
    return tmp_path, D, P, dead, live, logw


def test_dead_and_live_files(synthetic):
    base, D, P, dead, live, logw = synthetic
    np_ = D + P
    db = np.loadtxt(base / "syn_dead-birth.txt")            # theta, phi, logL, birth (what anesthetic reads)
    assert db.shape == (dead.shape[0], np_ + 2)
    assert np.allclose(db[:, :np_], dead[:, :np_], rtol=1e-14, atol=0)
    assert np.allclose(db[:, np_], dead[:, np_ + 1], rtol=1e-14) and np.allclose(db[:, np_ + 1], dead[:, np_], rtol=1e-14)
    dd = np.loadtxt(base / "syn_dead.txt")                  # logL, theta, phi
    assert np.allclose(dd[:, 0], dead[:, np_ + 1], rtol=1e-14) and np.allclose(dd[:, 1:], dead[:, :np_], rtol=1e-14)
    lb = np.loadtxt(base / "syn_phys_live-birth.txt")
    assert lb.shape == (live.shape[0], np_ + 2) and np.allclose(lb[:, np_], live[:, np_ + 1], rtol=1e-14)
    pl = np.loadtxt(base / "syn_phys_live.txt")             # theta, phi, logL
    assert pl.shape == (live.shape[0], np_ + 1)
    line = (base / "syn_dead-birth.txt").read_text().splitlines()[0]
    assert len(line) == 24 * (np_ + 2)                      # fixed-width Fortran records


def test_posterior_files(synthetic):
    base, D, P, dead, live, logw = synthetic
    np_ = D + P
    w = np.loadtxt(base / "syn.txt")                        # weight, -2 logL, theta, phi; maximum weight 1
    assert w.shape == (dead.shape[0], np_ + 2)
    assert np.isclose(w[:, 0].max(), 1.0) and np.allclose(w[:, 0], np.exp(logw - logw.max()), rtol=1e-13)
    assert np.allclose(w[:, 1], -2 * dead[:, np_ + 1], rtol=1e-14)
    eq = np.loadtxt(base / "syn_equal_weights.txt")
    assert np.all(eq[:, 0] == 1.0) and 0 < eq.shape[0] < dead.shape[0]
    # every equally weighted sample is one of the dead points; their number is about sum(w)/max(w)
    assert set(map(tuple, np.round(eq[:, 2:], 10))) <= set(map(tuple, np.round(dead[:, :np_], 10)))
    assert abs(eq.shape[0] - w[:, 0].sum()) < 5 * np.sqrt(w[:, 0].sum())


def test_stats_file_parses_like_the_reference_parser(synthetic):
    base, D, P, dead, live, logw = synthetic
    out = PolyChordOutput(str(base), "syn")
    assert out.logZ == -3.25 and out.logZerr == 0.125
    assert out.logZs == [-3.25] and out.ncluster == 1
    assert out.ndead == dead.shape[0] and out.nlive == live.shape[0] and out.nlike == 123456
    assert out.nposterior == dead.shape[0] and out.nequals == np.loadtxt(base / "syn_equal_weights.txt").shape[0]
    # <nlike> per update of nlive*|log compression| deaths, and per slice (read_write.F90:883-886)
    assert np.isclose(out.avnlike[0], 123456 / live.shape[0], atol=0.01)
    assert np.isclose(out.avnlikeslice[0], 123456 / live.shape[0] / 15, atol=0.01)
    # weighted posterior moments of theta and phi (read_write.F90:912-961)
    wts = np.exp(logw - logw.max()); wts /= wts.sum()
    mean = wts @ dead[:, :D + P]
    var = wts @ (dead[:, :D + P] - mean) ** 2
    assert np.allclose(out.means, mean, rtol=1e-10)
    assert np.allclose(out.sigmas, np.sqrt(var), rtol=1e-6)
    lines = (base / "syn.stats").read_text().split("\n")
    assert lines[0] == "Evidence estimates:" and lines[8].startswith("log(Z)       = ")
    assert lines[14].startswith("log(Z_1)     = ") and lines[14].endswith("(Still Active)")


def test_missing_base_dir_is_an_error(tmp_path):
    capi.set_option("errors_return", 1)
    try:
        rc = capi.write_files(tmp_path / "nope", "x", 1, 0, np.zeros((1, 3)), np.zeros(1), np.zeros((0, 3)), 0.0, 0.1, 1, 1)
    finally:
        capi.set_option("errors_return", 0)
    assert rc < 0


def test_posterior_files_with_boosted_samples(tmp_path):
    """boost_posterior: the promoted phantoms are written with the dead points, update by update (pc_write_files_boosted;
    update_posteriors, run_time_info.f90:1036-1061), and enter the weights' normalisation, the equally weighted file
    and the stats file's posterior count and means."""
    rng = np.random.default_rng(11)
    D, P, ndead, nb = 2, 1, 120, 300
    logL = np.sort(rng.uniform(-8, 0, ndead))
    dead = np.column_stack([rng.normal(0.5, 0.1, (ndead, D)), rng.uniform(0, 1, (ndead, P)), logL - 0.5, logL])
    dead_logw = -np.arange(ndead) / 30.0 + logL
    after = np.sort(rng.integers(1, 5, nb) * 30)               # removed at the updates after 30, 60, 90, 120 deaths
    bl = np.array([rng.uniform(logL[a - 30], logL[a - 1]) for a in after])
    boosted = np.column_stack([rng.normal(0.5, 0.1, (nb, D)), rng.uniform(0, 1, (nb, P)), bl - 0.5, bl])
    blogw = np.array([-(a - 15) / 30.0 for a in after]) + bl
    blogw[7] = dead_logw.max() + 1.0                           # a promoted phantom may carry the largest weight
    live = np.zeros((0, D + P + 2))
    n = capi.write_files(tmp_path, "b", D, P, dead, dead_logw, live, logZ=-1.0, logZerr=0.1, nlike=5000, num_repeats=4,
                         seed=3, flags=("stats", "posteriors", "equals"), boosted=(boosted, blogw, after))
    assert n == 3
    post = np.loadtxt(tmp_path / "b.txt")
    assert post.shape == (ndead + nb, 2 + D + P)
    allw = np.concatenate([dead_logw, blogw])
    assert np.isclose(post[:, 0].max(), 1.0)
    assert np.allclose(np.sort(post[:, 0]), np.sort(np.exp(allw - allw.max())), rtol=1e-12)
    # order: the dead points of an update, then the phantoms it removed
    is_dead = np.isin(np.round(post[:, 1], 9), np.round(-2 * logL, 9))
    k = 0
    for upd in (30, 60, 90, 120):
        m = int(np.sum(after == upd))
        assert is_dead[k:k + 30].all() and not is_dead[k + 30:k + 30 + m].any()
        assert np.allclose(post[k:k + 30, 1], -2 * logL[upd - 30:upd], rtol=1e-13)
        k += 30 + m
    eq = np.loadtxt(tmp_path / "b_equal_weights.txt")
    assert np.all(eq[:, 0] == 1.0)
    assert abs(eq.shape[0] - post[:, 0].sum()) < 5 * np.sqrt(post[:, 0].sum())
    stats = (tmp_path / "b.stats").read_text()
    assert f"nposterior: {ndead + nb:8d}" in stats
    # the same call without the samples writes the dead points only
    capi.write_files(tmp_path, "p", D, P, dead, dead_logw, live, logZ=-1.0, logZerr=0.1, nlike=5000, num_repeats=4,
                     seed=3, flags=("posteriors",))
    assert np.loadtxt(tmp_path / "p.txt").shape[0] == ndead
    # samples out of update order are refused
    capi.set_option("errors_return", 1)
    try:
        assert capi.write_files(tmp_path, "x", D, P, dead, dead_logw, live, logZ=-1.0, logZerr=0.1, nlike=5000, num_repeats=4,
                                seed=3, flags=("posteriors",), boosted=(boosted, blogw, after[::-1])) < 0
    finally:
        capi.set_option("errors_return", 0)


def test_prior_file(tmp_path):
    """<root>_prior.txt (write_prior_file, read_write.F90:721-752): the points drawn from the prior, [1, -2 logL, theta,
    phi], wherever they are by now (dead or still live)."""
    rng = np.random.default_rng(5)
    D, P = 3, 1

    def rows(n, nprior):
        logL = np.sort(rng.uniform(-9, 0, n))
        birth = np.where(np.arange(n) < nprior, -1e30, logL - 0.3)
        return np.column_stack([rng.uniform(0, 1, (n, D)), rng.uniform(0, 1, (n, P)), birth, logL])
    dead, live = rows(80, 30), rows(40, 10)
    n = capi.write_files(tmp_path, "pr", D, P, dead, dead[:, -1] - 1.0, live, logZ=-2.0, logZerr=0.2, nlike=999,
                         num_repeats=6, flags=("prior",))
    assert n == 1
    pr = np.loadtxt(tmp_path / "pr_prior.txt")
    assert pr.shape == (40, 2 + D + P) and np.all(pr[:, 0] == 1.0)
    want = np.vstack([dead[:30], live[:10]])
    assert np.allclose(pr[:, 1], -2 * want[:, -1], rtol=1e-14) and np.allclose(pr[:, 2:], want[:, :D + P], rtol=1e-14)
    assert len((tmp_path / "pr_prior.txt").read_text().splitlines()[0]) == 24 * (2 + D + P)
