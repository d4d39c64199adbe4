"""Prior transforms, same classes and call semantics as the reference's pypolychord/priors.py:5-47.

Every class is a plain callable `theta = prior(cube)` (numpy in, numpy out), as in the reference.
`UniformPrior` additionally has a device form (priors.f90:40-55, uniform_htp): when it is passed to
`run()` the whole sampling loop runs on the GPU.  The other transforms are not affine, so they need the
generic host-callback path (not built yet): `run()` says so.
"""
import numpy
from scipy.special import erfinv


class UniformPrior:
    def __init__(self, a, b):
        self.a = a
        self.b = b

    def __call__(self, x):
        return self.a + (self.b - self.a) * x

    def device_params(self, nDims):
        """lo[D], hi[D] for pc_register_device_prior (PC_PRIOR_UNIFORM)."""
        lo = numpy.broadcast_to(numpy.asarray(self.a, dtype=float), (nDims,))
        hi = numpy.broadcast_to(numpy.asarray(self.b, dtype=float), (nDims,))
        return numpy.concatenate([lo, hi])


class GaussianPrior:
    def __init__(self, mu, sigma):
        self.mu = mu
        self.sigma = sigma

    def __call__(self, x):
        return self.mu + self.sigma * numpy.sqrt(2) * erfinv(2 * x - 1)


class LogUniformPrior(UniformPrior):
    def __call__(self, x):
        return self.a * (self.b / self.a) ** x

    device_params = None


def forced_indentifiability_transform(x):
    N = len(x)
    t = numpy.zeros(N)
    t[N - 1] = x[N - 1] ** (1. / N)
    for n in range(N - 2, -1, -1):
        t[n] = x[n] ** (1. / (n + 1)) * t[n + 1]
    return t


class SortedUniformPrior(UniformPrior):
    def __call__(self, x):
        t = forced_indentifiability_transform(x)
        return super(SortedUniformPrior, self).__call__(t)

    device_params = None


class LogSortedUniformPrior(LogUniformPrior):
    def __call__(self, x):
        t = forced_indentifiability_transform(x)
        return super(LogSortedUniformPrior, self).__call__(t)
