"""Results of a run.

The reference returns `anesthetic.read_chains(base_dir/file_root)` from `run()` (polychord.py:639-646) and a
`PolyChordOutput` parsed from `<root>.stats` from `run_polychord()` (output.py:57-99).  The engine writes those
files in the reference's formats (csrc/pc_files.cpp), so both work: `run()` hands the chains to anesthetic when it
is importable, and otherwise returns the small in-memory `NestedSamplesLite` built from the final dumper call
(nested_sampling.F90:546-590) -- the dead points with birth contours and posterior weights, i.e. the content of
`<root>_dead-birth.txt`.
"""
import os
import re

import numpy as np


class PolyChordOutput:
    """`<root>.stats` parsed by the same fixed line offsets the reference uses (output.py:57-99), plus
    readers for the other files the engine writes."""

    def __init__(self, base_dir, file_root):
        self.base_dir, self.file_root = base_dir, file_root
        with open(self.root + '.stats') as f:
            lines = f.read().split('\n')
        line = lines[8]                                   # "log(Z)       = mu +/- sigma"
        self.logZ, self.logZerr = float(line.split()[2]), float(line.split()[4])
        i = 14                                            # first "log(Z_p)" line
        self.logZs, self.logZerrs = [], []
        while lines[i][:5] == 'log(Z':
            rhs = re.findall(r'=(.*)', lines[i])[0].split()
            self.logZs.append(float(rhs[0]))
            self.logZerrs.append(float(rhs[2]))
            i += 1
        i += 5                                            # blank, blank, title, rule, blank
        self.ncluster = len(self.logZs)
        i += 1                                            # " ncluster:" line
        self.nposterior = int(lines[i].split()[1]); i += 1
        self.nequals = int(lines[i].split()[1]); i += 1
        self.ndead = int(lines[i].split()[1]); i += 1
        self.nlive = int(lines[i].split()[1]); i += 1
        try:
            self.nlike = int(lines[i].split()[1])
        except ValueError:                                # "********" when the count overflows I8
            self.nlike = None
        i += 1
        tok = lines[i].split()
        j = tok.index('(')
        self.avnlike = [float(x) for x in tok[1:j]]
        self.avnlikeslice = [float(x) for x in tok[j + 1:-3]]
        # "Dim No.       Mean        Sigma" table, when posteriors were requested
        self.means, self.sigmas = [], []
        for ln in lines[i + 1:]:
            m = re.match(r'\s*(\d+)\s+(\S+)\s+\+/-\s+(\S+)', ln)
            if m:
                self.means.append(float(m.group(2)))
                self.sigmas.append(float(m.group(3)))

    @property
    def root(self):
        return os.path.join(self.base_dir, self.file_root)

    def cluster_root(self, i):
        return os.path.join(self.base_dir, 'clusters', '%s_%i' % (self.file_root, i))

    @property
    def paramnames_file(self):
        return self.root + '.paramnames'

    def make_paramnames_files(self, paramnames):
        make_paramnames_file(paramnames, self.paramnames_file)

    def dead_birth(self):
        """rows [theta, phi, logL, logL_birth] of `<root>_dead-birth.txt` (what anesthetic reads)"""
        return np.atleast_2d(np.loadtxt(self.root + '_dead-birth.txt'))

    def weighted_posterior(self):
        """rows [weight, -2 logL, theta, phi] of `<root>.txt`"""
        return np.atleast_2d(np.loadtxt(self.root + '.txt'))

    def equal_weights(self):
        return np.atleast_2d(np.loadtxt(self.root + '_equal_weights.txt'))

    def __str__(self):
        s = 'Global evidence:\n  log(Z)       = %s +/- %s\n' % (self.logZ, self.logZerr)
        for k, (z, e) in enumerate(zip(self.logZs, self.logZerrs)):
            s += '  log(Z_%i)  = %s +/- %s\n' % (k + 1, z, e)
        s += 'ncluster: %i\nnposterior: %i\nnequals: %i\nndead: %i\nnlive: %i\nnlike: %s\n' % (
            self.ncluster, self.nposterior, self.nequals, self.ndead, self.nlive, self.nlike)
        return s


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
