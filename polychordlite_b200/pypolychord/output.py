"""Results of a run.

The reference returns `anesthetic.read_chains(base_dir/file_root)` (polychord.py:639-646) or a
`PolyChordOutput` parsed from `<root>.stats` (output.py:57-99).  File output is SURVEY.md section 8 row f1
(not built yet) and anesthetic is not installed here, so `run()` returns this small in-memory object built
from the final dumper call (nested_sampling.F90:546-590): the dead points with birth contours and posterior
weights, which is what anesthetic reads from `<root>_dead-birth.txt`.
"""
import numpy as np


class NestedSamplesLite:
    def __init__(self, dead, logweights, logZ, logZerr, nDims, nDerived, info=None):
        dead = np.asarray(dead)
        self.nDims, self.nDerived = nDims, nDerived
        self.theta = dead[:, :nDims]
        self.phi = dead[:, nDims:nDims + nDerived]
        self.logL_birth = dead[:, nDims + nDerived]
        self.logL = dead[:, nDims + nDerived + 1]
        self.logweights = np.asarray(logweights)   # normalised posterior log-weights
        self.logZ, self.logZerr = logZ, logZerr
        self.info = info or {}

    @property
    def ndead(self):
        return self.theta.shape[0]

    @property
    def weights(self):
        w = np.exp(self.logweights - self.logweights.max())
        return w / w.sum()

    def mean(self):
        return self.weights @ self.theta

    def std(self):
        m = self.mean()
        return np.sqrt(self.weights @ (self.theta - m) ** 2)

    def equals(self, other):
        """Same role as pandas' .equals in the reference's tests (test_run_pypolychord.py:77-119)."""
        return (self.theta.shape == other.theta.shape and np.array_equal(self.theta, other.theta)
                and np.array_equal(self.logL, other.logL) and np.array_equal(self.logL_birth, other.logL_birth)
                and np.array_equal(self.logweights, other.logweights))


def make_paramnames_file(paramnames, filename):
    """output.py:131-143: `<name>   <latex>` per line."""
    with open(filename, 'w') as f:
        for name, latex in paramnames:
            f.write('%s   %s\n' % (name, latex))
