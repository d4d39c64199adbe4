"""Host-side mirror of the reference's `pypolychord` package for the hot path.

Same entry points, keyword arguments, defaults and error behaviour as
/root/reference/pypolychord/{polychord,settings,priors,output}.py, driving the B200 engine through the
C ABI (`polychord_c_interface`, include/polychord_b200.h) instead of the Fortran library:

    from polychordlite_b200 import pypolychord
    from polychordlite_b200.pypolychord.builtin import Gaussian
    from polychordlite_b200.pypolychord.priors import UniformPrior
    samples = pypolychord.run(Gaussian(mu=0.5, sigma=0.1, nDerived=2), nDims=20, nDerived=2,
                              prior=UniformPrior(0, 1), nlive=1000, num_repeats=40, seed=1)
"""
from .polychord import run, run_polychord, default_prior, default_dumper  # noqa: F401
from .settings import PolyChordSettings  # noqa: F401
from .output import NestedSamplesLite, PolyChordOutput  # noqa: F401
from . import priors, builtin  # noqa: F401

__version__ = "1.22.2+b200"
