"""The reference's own test file, tests/test_run_pypolychord.py, restated against this package: same likelihood
(a Python callable returning (logL, [r2])), same prior object, same settings objects, same assertions.  What differs,
and why: anesthetic is not installed in this image, so `run()` returns the in-memory NestedSamplesLite (its
`.equals` plays the role of pandas' in test_seed / test_no_derived); `cube_samples` (resume-file injection, row f3) is outside this round's scope."""
import numpy as np
import pytest

from polychordlite_b200 import pypolychord
from polychordlite_b200.pypolychord.priors import UniformPrior
from polychordlite_b200.pypolychord.settings import PolyChordSettings

pytestmark = pytest.mark.gpu


def gaussian_likelihood(theta):
    sigma = 0.1
    nDims = len(theta)
    r2 = sum(theta ** 2)
    logL = -np.log(2 * np.pi * sigma * sigma) * nDims / 2.0
    logL += -r2 / 2 / sigma / sigma
    return logL, [r2]


uniform_prior = UniformPrior(-1, 1)


def test_run_polychord(gpu, tmp_path):                     # test_run_pypolychord.py:40-59
    settings = PolyChordSettings(4, 1)
    settings.file_root = 'settings'
    settings.nlive = 200
    settings.read_resume = False
    settings.feedback = 0
    settings.base_dir = str(tmp_path)
    last = []

    def dumper(live, dead, logweights, logZ, logZerr):
        last.append(dead[-1].copy())

    output = pypolychord.run_polychord(gaussian_likelihood, 4, 1, settings, uniform_prior, dumper)
    paramnames = [('p%i' % i, r'\theta_%i' % i) for i in range(4)] + [('r*', 'r')]
    output.make_paramnames_files(paramnames)
    assert last and (tmp_path / 'settings.paramnames').exists() and (tmp_path / 'settings.stats').exists()
    assert abs(output.logZ - (-4 * np.log(2))) < 0.6       # the numeric check the reference lacks


def test_run(gpu, tmp_path):                               # :62-74
    paramnames = [('p%i' % i, r'\theta_%i' % i) for i in range(4)] + [('r*', 'r')]
    ns = pypolychord.run(gaussian_likelihood, 4, nDerived=1, prior=uniform_prior, paramnames=paramnames,
                         read_resume=False, base_dir=str(tmp_path), feedback=0)
    assert ns.ndead > 500 and (tmp_path / 'test_dead-birth.txt').exists()
    assert np.allclose(ns.phi[:, 0], np.sum(ns.theta ** 2, axis=1))


@pytest.mark.parametrize("seed", [-1, 0, 1, 2])
def test_seed(gpu, tmp_path, seed):                        # :77-90
    kw = dict(nDerived=1, prior=uniform_prior, read_resume=False, seed=seed, base_dir=str(tmp_path), feedback=0)
    ns0 = pypolychord.run(gaussian_likelihood, 4, **kw)
    ns1 = pypolychord.run(gaussian_likelihood, 4, **kw)
    assert ns0.equals(ns1) != (seed < 0)


def test_no_derived(gpu, tmp_path):                        # :93-119
    def no_derived_gaussian_likelihood(theta):
        return gaussian_likelihood(theta)[0]

    kw = dict(prior=uniform_prior, read_resume=False, seed=1, base_dir=str(tmp_path), feedback=0)
    ns0 = pypolychord.run(no_derived_gaussian_likelihood, 4, **kw)
    ns1 = pypolychord.run(gaussian_likelihood, 4, nDerived=1, **kw)
    assert ns0.equals(ns1)                                 # .equals ignores the derived column, like drop(columns='r')


def test_grade_dims(gpu, tmp_path):                        # :122-130
    with pytest.raises(ValueError):
        pypolychord.run(gaussian_likelihood, 5, nDerived=1, prior=uniform_prior, read_resume=False, grade_dims=[1, 3],
                        base_dir=str(tmp_path))
    ns = pypolychord.run(gaussian_likelihood, 4, nDerived=1, prior=uniform_prior, read_resume=False, grade_dims=[1, 3],
                         base_dir=str(tmp_path), feedback=0)
    # two grades with equal fractions: 20 slice steps in all 4 dimensions + 20 in the 3 fast ones per chain
    assert ns.info["nslices"] == 40 * ns.info["nchains"]
    assert abs(ns.logZ - (-4 * np.log(2))) < 0.8
