"""PolyChordSettings: same attributes, defaults and errors as the reference's pypolychord/settings.py:176-222.

The reference's constructor pops one keyword per attribute; here the attributes and their defaults are one table
(a default that depends on nDims is a callable), applied in a loop."""
import os

import numpy

# attribute -> default, or a function of nDims (the table of settings.py:176-218; documentation there, :10-170)
_DEFAULTS = (
    ('nlive', lambda nDims: nDims * 25),
    ('num_repeats', lambda nDims: nDims * 5),
    ('nprior', -1), ('nfail', -1),
    ('do_clustering', True),
    ('feedback', 1),
    ('precision_criterion', 0.001),
    ('logzero', -1e30),
    ('max_ndead', -1),
    ('boost_posterior', 0.0),
    ('posteriors', True), ('equals', True), ('cluster_posteriors', True),
    ('write_resume', True), ('write_paramnames', False), ('read_resume', True),
    ('write_stats', True), ('write_live', True), ('write_dead', True), ('write_prior', True),
    ('maximise', False),
    ('compression_factor', float(numpy.exp(-1))),
    ('synchronous', True),
    ('base_dir', 'chains'), ('file_root', 'test'),
    ('seed', -1),
    ('grade_dims', lambda nDims: [nDims]),
    ('grade_frac', None),            # one entry per grade, filled in below
    ('nlives', dict),
    ('cube_samples', None),
)


class PolyChordSettings:
    def __init__(self, nDims, nDerived, **kwargs):
        for name, default in _DEFAULTS:
            if name in kwargs:
                value = kwargs.pop(name)
            elif default is dict:
                value = {}
            else:
                value = default(nDims) if callable(default) else default
            setattr(self, name, value)
        self.grade_dims = list(self.grade_dims)
        self.grade_frac = [1.0] * len(self.grade_dims) if self.grade_frac is None else list(self.grade_frac)

        if kwargs:
            raise TypeError('Unexpected **kwargs in Contours constructor: %r' % kwargs)   # (the reference's wording)
        if sum(self.grade_dims) != nDims:
            raise ValueError('grade_dims must sum to the total dimensionality:'
                             'sum(%s) /= %i' % (self.grade_dims, nDims))

    @property
    def cluster_dir(self):
        return os.path.join(self.base_dir, 'clusters')
