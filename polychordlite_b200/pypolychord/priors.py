"""Prior transforms with the public surface of the reference's pypolychord/priors.py (:5-47): the same class names,
the same inheritance (`isinstance(p, UniformPrior)` holds for the log and sorted variants) and `theta = prior(cube)`.

Built here from two pieces -- a separable map applied coordinate by coordinate, optionally preceded by the ordering
map of the unit cube -- instead of one method per class.  Only the plain `UniformPrior` is affine, so only it has a
device form (`device_params`, consumed by pc_register_device_prior; priors.f90:40-55 uniform_htp) and keeps the whole
sampling loop on the GPU; every other prior runs through the host-callback path.
"""
import numpy as np
from scipy.special import ndtri


def forced_indentifiability_transform(x):
    """Unit cube -> its ordered corner (t[0] <= t[1] <= ...), volume preserving: the largest coordinate is the
    maximum of N uniforms, each one below it the maximum of the remaining ones within the bound just set
    (priors.f90:242-264 sort_hypercube; the spelling of the name is the reference's)."""
    x = np.asarray(x, dtype=float)
    n = x.shape[0]
    roots = x ** (1.0 / np.arange(1, n + 1))
    # t[k] = roots[k] * roots[k+1] * ... * roots[n-1], accumulated from the top down
    return np.cumprod(roots[::-1])[::-1]


class _Separable:
    """theta_i = f(cube_i; p, q) with an optional ordering of the cube first."""
    _ordered = False

    def __init__(self, p, q):
        self._p, self._q = p, q

    def _map(self, u):
        raise NotImplementedError

    def __call__(self, x):
        u = forced_indentifiability_transform(x) if self._ordered else x
        return self._map(u)

    device_params = None   # no device form unless a subclass provides one


class UniformPrior(_Separable):
    """Flat between a and b."""

    def __init__(self, a, b):
        super().__init__(a, b)

    a = property(lambda self: self._p)
    b = property(lambda self: self._q)

    def _map(self, u):
        return self._p + (self._q - self._p) * u

    def device_params(self, nDims):
        """lo[D], hi[D] for pc_register_device_prior (PC_PRIOR_UNIFORM)."""
        ends = [np.broadcast_to(np.asarray(v, dtype=float), (nDims,)) for v in (self._p, self._q)]
        return np.concatenate(ends)


class GaussianPrior(_Separable):
    """Normal with mean mu and standard deviation sigma (inverse normal CDF of the cube coordinate)."""

    def __init__(self, mu, sigma):
        super().__init__(mu, sigma)

    mu = property(lambda self: self._p)
    sigma = property(lambda self: self._q)

    def _map(self, u):
        return self._p + self._q * ndtri(u)


class LogUniformPrior(UniformPrior):
    """Flat in log(theta) between a and b."""
    device_params = None

    def _map(self, u):
        return self._p * np.power(self._q / self._p, u)


class SortedUniformPrior(UniformPrior):
    """Uniform with theta_1 <= theta_2 <= ... enforced."""
    _ordered = True
    device_params = None


class LogSortedUniformPrior(LogUniformPrior):
    """Log-uniform with theta_1 <= theta_2 <= ... enforced."""
    _ordered = True
