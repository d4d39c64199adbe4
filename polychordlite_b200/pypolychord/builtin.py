"""The reference's built-in analytic likelihoods (likelihoods/examples/*.f90) as Python objects with a
device form.  Passing one of these to `run()` keeps the whole sampling loop on the GPU (the ABI only carries
host function pointers, so the engine recognises them through pc_register_device_likelihood, see
include/polychord_b200.h).  They are ordinary callables too: `logL, phi = Gaussian(...)(theta)` evaluates the
library's host callback, the same formula as the device code.
"""
import ctypes as C

import numpy as np

from .. import _capi


class _DeviceLikelihood:
    kind = None
    host_symbol = None

    def __init__(self, nDerived=0):
        self.nDerived = int(nDerived)

    def device_params(self, nDims):
        return np.zeros(0)

    def host_fn(self):
        return getattr(_capi.lib(), self.host_symbol)

    def register(self, nDims):
        L = _capi.lib()
        p = np.ascontiguousarray(self.device_params(nDims), dtype=np.float64)
        fn = C.cast(self.host_fn(), _capi.LL_CB)
        rc = L.pc_register_device_likelihood(fn, self.kind, p.ctypes.data_as(C.POINTER(C.c_double)), p.size)
        if rc != 0:
            raise RuntimeError("pc_register_device_likelihood failed")
        return fn

    def __call__(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        self.register(theta.size)
        phi = np.zeros(max(self.nDerived, 1))
        fn = self.host_fn()
        fn.restype = C.c_double
        logL = fn(theta.ctypes.data_as(C.POINTER(C.c_double)), theta.size,
                  phi.ctypes.data_as(C.POINTER(C.c_double)), self.nDerived)
        return (logL, phi[:self.nDerived]) if self.nDerived else logL


class Gaussian(_DeviceLikelihood):
    """likelihoods/examples/gaussian.f90:12-41: independent Gaussians; derived: |theta-mu| and log(r^D V_D)."""
    kind = 0
    host_symbol = "pc_gaussian_loglikelihood"

    def __init__(self, mu=0.5, sigma=0.1, nDerived=0):
        super().__init__(nDerived)
        self.mu, self.sigma = mu, sigma

    def device_params(self, nDims):
        if np.ndim(self.mu) == 0 and np.ndim(self.sigma) == 0:
            return np.array([float(self.mu), float(self.sigma)])   # the engine's dimension-free form: every dimension alike
        mu = np.broadcast_to(np.asarray(self.mu, dtype=float), (nDims,))
        sg = np.broadcast_to(np.asarray(self.sigma, dtype=float), (nDims,))
        return np.concatenate([mu, sg])


class Rastrigin(_DeviceLikelihood):
    """likelihoods/examples/rastrigin.f90:20-35 (normalised on [-5.12, 5.12]^D)."""
    kind = 1
    host_symbol = "pc_rastrigin_loglikelihood"


class CorrelatedGaussian(_DeviceLikelihood):
    """likelihoods/examples/random_gaussian.f90 / utils.F90:1028-1048 log_gauss: mean mu, inverse covariance
    invcov (D x D), logdet = log det(covariance)."""
    kind = 2
    host_symbol = "pc_corr_gaussian_loglikelihood"

    def __init__(self, mu, invcov, logdet, nDerived=0):
        super().__init__(nDerived)
        self.mu, self.invcov, self.logdet = np.asarray(mu, float), np.asarray(invcov, float), float(logdet)

    def device_params(self, nDims):
        return np.concatenate([np.broadcast_to(self.mu, (nDims,)), self.invcov.ravel(order="F"), [self.logdet]])
