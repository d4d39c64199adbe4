"""PolyChordSettings: same attributes, defaults and errors as the reference's pypolychord/settings.py:176-222."""
import os

import numpy


class PolyChordSettings:
    def __init__(self, nDims, nDerived, **kwargs):
        self.nlive = kwargs.pop('nlive', nDims * 25)
        self.num_repeats = kwargs.pop('num_repeats', nDims * 5)
        self.nprior = kwargs.pop('nprior', -1)
        self.nfail = kwargs.pop('nfail', -1)
        self.do_clustering = kwargs.pop('do_clustering', True)
        self.feedback = kwargs.pop('feedback', 1)
        self.precision_criterion = kwargs.pop('precision_criterion', 0.001)
        self.logzero = kwargs.pop('logzero', -1e30)
        self.max_ndead = kwargs.pop('max_ndead', -1)
        self.boost_posterior = kwargs.pop('boost_posterior', 0.0)
        self.posteriors = kwargs.pop('posteriors', True)
        self.equals = kwargs.pop('equals', True)
        self.cluster_posteriors = kwargs.pop('cluster_posteriors', True)
        self.write_resume = kwargs.pop('write_resume', True)
        self.write_paramnames = kwargs.pop('write_paramnames', False)
        self.read_resume = kwargs.pop('read_resume', True)
        self.write_stats = kwargs.pop('write_stats', True)
        self.write_live = kwargs.pop('write_live', True)
        self.write_dead = kwargs.pop('write_dead', True)
        self.write_prior = kwargs.pop('write_prior', True)
        self.maximise = kwargs.pop('maximise', False)
        self.compression_factor = kwargs.pop('compression_factor', numpy.exp(-1))
        self.synchronous = kwargs.pop('synchronous', True)
        self.base_dir = kwargs.pop('base_dir', 'chains')
        self.file_root = kwargs.pop('file_root', 'test')
        self.seed = kwargs.pop('seed', -1)
        self.grade_dims = list(kwargs.pop('grade_dims', [nDims]))
        self.grade_frac = list(kwargs.pop('grade_frac', [1.0] * len(self.grade_dims)))
        self.nlives = kwargs.pop('nlives', {})
        self.cube_samples = kwargs.pop('cube_samples', None)

        if kwargs:
            raise TypeError('Unexpected **kwargs in Contours constructor: %r' % kwargs)

        if sum(self.grade_dims) != nDims:
            raise ValueError('grade_dims must sum to the total dimensionality:'
                             'sum(%s) /= %i' % (self.grade_dims, nDims))

    @property
    def cluster_dir(self):
        return os.path.join(self.base_dir, 'clusters')
