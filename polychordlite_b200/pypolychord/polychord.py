"""run() / run_polychord(): the reference's Python entry points (pypolychord/polychord.py:16 and :221)
on top of the B200 engine's C ABI.

Differences from the reference, all forced by scope:
  * likelihoods with a device form (`pypolychord.builtin.*`) with the default unit-cube prior or a `UniformPrior`
    run the whole sampling loop on the GPU; any other Python callable (likelihood or prior) takes the engine's
    host-callback path (row f2): the chains advance in lock step on the device and call back once per trial point;
  * anesthetic is not installed in this image: `run()` returns `anesthetic.read_chains(...)` when it can be
    imported and an in-memory `NestedSamplesLite` built from the final dumper call otherwise; `run_polychord()`
    returns a `PolyChordOutput` parsed from the `<root>.stats` file the engine wrote (row f1).
Everything else -- keyword names, defaults, TypeError on unknown keywords, ValueError when grade_dims does
not sum to nDims, creation of base_dir/cluster_dir, the paramnames file, the dumper signature -- follows the
reference line by line (polychord.py:520-595).
"""
import ctypes as C
from pathlib import Path

import numpy as np

import os

from .. import _capi
from . import builtin as _builtin

try:  # the CPython extension built by polychordlite_b200/_build.py (reference: pypolychord/_pypolychord.cpp)
    from . import _pypolychord as _shim
    if os.environ.get("PC_PY_BINDING") == "ctypes":
        _shim = None
except ImportError:  # not built: the same C ABI bound with ctypes
    _shim = None
from .output import NestedSamplesLite, PolyChordOutput, make_paramnames_file
from .priors import UniformPrior


def default_prior(cube):
    return cube.copy()


def default_dumper(live, dead, logweights, logZ, logZerr):
    pass


_ARGTYPES = ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_bool, C.c_int, C.c_double,
              C.c_double, C.c_int, C.c_double] + [C.c_bool] * 11 +
             [C.c_double, C.c_bool, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_double),
              C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)])


def _view(ptr, shape):
    """zero-copy float64 view of engine-owned memory (what the reference's shim does with
    PyArray_SimpleNewFromData, _pypolychord.cpp:88-94)"""
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape)
    buf = (C.c_double * n).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=np.float64).reshape(shape)


def run(loglikelihood, nDims, **kwargs):
    paramnames = kwargs.pop('paramnames', None)
    default_kwargs = {
        'nDerived': 0, 'prior': default_prior, 'dumper': default_dumper, 'nlive': nDims * 25,
        'num_repeats': nDims * 5, 'nprior': -1, 'nfail': -1, 'do_clustering': True, 'feedback': 1,
        'precision_criterion': 0.001, 'logzero': -1e30, 'max_ndead': -1, 'boost_posterior': 0.0,
        'posteriors': True, 'equals': True, 'cluster_posteriors': True, 'write_resume': True,
        'write_paramnames': False, 'read_resume': True, 'write_stats': True, 'write_live': True,
        'write_dead': True, 'write_prior': True, 'maximise': False, 'compression_factor': np.exp(-1),
        'synchronous': True, 'base_dir': 'chains', 'file_root': 'test', 'cluster_dir': 'clusters',
        'grade_dims': [nDims], 'nlives': {}, 'seed': -1, 'cube_samples': None,
    }
    default_kwargs['grade_frac'] = ([1.0] * len(default_kwargs['grade_dims']) if 'grade_dims' not in kwargs
                                    else [1.0] * len(kwargs['grade_dims']))
    legacy = kwargs.pop('_legacy_output', False)
    if not kwargs.keys() <= default_kwargs.keys():
        raise TypeError(f"{__name__} got unknown keyword arguments: {kwargs.keys() - default_kwargs.keys()}")
    default_kwargs.update(kwargs)
    kwargs = default_kwargs
    kwargs['_legacy_output'] = legacy

    (Path(kwargs['base_dir']) / kwargs['cluster_dir']).mkdir(parents=True, exist_ok=True)
    if paramnames is not None:
        make_paramnames_file(paramnames, Path(kwargs['base_dir']) / (kwargs['file_root'] + ".paramnames"))

    kwargs['grade_dims'] = [int(d) for d in list(kwargs['grade_dims'])]
    if sum(kwargs['grade_dims']) != nDims:
        raise ValueError(f"grade_dims ({sum(kwargs['grade_dims'])}) must sum to nDims ({nDims})")
    kwargs['nlives'] = {float(logL): int(nlive) for logL, nlive in kwargs['nlives'].items()}

    L = _capi.lib()
    nDerived = int(kwargs['nDerived'])
    prior = kwargs['prior']
    user_dumper = kwargs['dumper']
    last = {}

    def on_final_dump(dead, logweights, logZ, logZerr):
        last.update(dead=dead.copy(), logweights=logweights.copy(), logZ=logZ, logZerr=logZerr)

    # the reference convention for fatal configuration errors is a banner and exit(1) (abort.F90:19-29);
    # under Python report them as exceptions instead
    old_err = _capi.get_option("errors_return")
    _capi.set_option("errors_return", 1)
    try:
        if kwargs.get('cube_samples') is not None:
            # polychord.py:576-579: the run starts from the caller's live points; any number of them (the dynamic-nlive
            # schedule absorbs the difference from nlive, as in the reference).  The points win over an existing resume file.
            cs = np.asarray(kwargs['cube_samples'], dtype=np.float64)
            if cs.ndim != 2 or cs.shape[1] != nDims or cs.shape[0] < 2:
                raise ValueError("cube_samples must have shape (npoints, nDims)")
            _capi.set_initial_live(cs)
        if _shim is not None:
            _run_through_shim(loglikelihood, prior, user_dumper, on_final_dump, kwargs, nDims, nDerived)
        else:
            _run_through_ctypes(L, loglikelihood, prior, user_dumper, on_final_dump, kwargs, nDims, nDerived)
    finally:
        _capi.set_option("errors_return", old_err)
    info = _capi.last_run_info()
    if info.status != 0 or not last:
        raise RuntimeError(f"polychord_c_interface failed (status {info.status}); see the message on stderr")
    lite = NestedSamplesLite(last['dead'], last['logweights'], last['logZ'], last['logZerr'], nDims, nDerived,
                             info=info.as_dict())
    if kwargs.get('_legacy_output'):
        return lite
    try:  # polychord.py:639-646: the chains as anesthetic reads them from the files the engine wrote
        import anesthetic
        if kwargs['write_dead']:
            return anesthetic.read_chains(str(Path(kwargs['base_dir']) / kwargs['file_root']))
    except ImportError:
        pass
    return lite


class _CAddress:
    """A callable with a device form, as the `_pypolychord` extension recognises it: `_pc_c_callback(nDims)` returns the
    address of a C callback the engine knows (pc_register_device_likelihood / pc_register_device_prior)."""

    def __init__(self, address_of):
        self._address_of = address_of

    def _pc_c_callback(self, nDims):
        return self._address_of(nDims)

    def __call__(self, *a):  # never called: the engine runs the device form
        raise RuntimeError("device-resident callback called on the host")


def _device_prior(prior, nDims, L):
    """The C address of a prior with a device form (unit cube, UniformPrior), or None."""
    if prior is default_prior:
        return C.cast(L.pc_unit_prior, C.c_void_p).value
    if isinstance(prior, UniformPrior) and getattr(prior, 'device_params', None) is not None:
        fn = C.cast(L.pc_uniform_prior, _capi.PRIOR_CB)
        pp = np.ascontiguousarray(prior.device_params(nDims), dtype=np.float64)
        if L.pc_register_device_prior(fn, 0, pp.ctypes.data_as(C.POINTER(C.c_double)), pp.size) != 0:
            raise RuntimeError("pc_register_device_prior failed")
        return C.cast(fn, C.c_void_p).value
    return None


def _run_through_shim(loglikelihood, prior, user_dumper, on_final_dump, kwargs, nDims, nDerived):
    """polychord.py:581-634: wrap the callables and hand the 34 positional arguments to `_pypolychord.run`.  A Python
    exception raised inside a callback unwinds through the engine and is re-raised by the extension."""
    L = _capi.lib()
    if isinstance(loglikelihood, _builtin._DeviceLikelihood):
        like = _CAddress(lambda nd: C.cast(loglikelihood.register(nd), C.c_void_p).value)
    else:
        def like(theta, phi):  # polychord.py:581-587 wrap_loglikelihood
            res = loglikelihood(theta)
            if isinstance(res, tuple):
                logL, derived = res
                if phi.size:
                    phi[:] = derived
            else:
                logL = res
            return float(logL)
    addr = _device_prior(prior, nDims, L)
    if addr is not None:
        prior_cb = _CAddress(lambda nd: addr)
    else:
        def prior_cb(cube, theta):  # polychord.py:589-590 wrap_prior
            theta[:] = prior(cube)

    def dumper(live, dead, logweights, logZ, logZerr):
        if live.shape[0] == 0:  # the final call: every point is dead (nested_sampling.F90:392)
            on_final_dump(dead, logweights, logZ, logZerr)
        user_dumper(live, dead, logweights, logZ, logZerr)

    _shim.run(like, prior_cb, dumper, nDims, nDerived, int(kwargs['nlive']), int(kwargs['num_repeats']),
              int(kwargs['nprior']), int(kwargs['nfail']), bool(kwargs['do_clustering']), int(kwargs['feedback']),
              float(kwargs['precision_criterion']), float(kwargs['logzero']), int(kwargs['max_ndead']),
              float(kwargs['boost_posterior']), bool(kwargs['posteriors']), bool(kwargs['equals']),
              bool(kwargs['cluster_posteriors']), bool(kwargs['write_resume']), bool(kwargs['write_paramnames']),
              bool(kwargs['read_resume']), bool(kwargs['write_stats']), bool(kwargs['write_live']),
              bool(kwargs['write_dead']), bool(kwargs['write_prior']), bool(kwargs['maximise']),
              float(kwargs['compression_factor']), bool(kwargs['synchronous']), str(kwargs['base_dir']),
              str(kwargs['file_root']), [float(f) for f in kwargs['grade_frac']], kwargs['grade_dims'], kwargs['nlives'],
              int(kwargs['seed']))


def _run_through_ctypes(L, loglikelihood, prior, user_dumper, on_final_dump, kwargs, nDims, nDerived):
    """The same call bound with ctypes (used when the `_pypolychord` extension has not been built)."""
    # ---- loglikelihood / prior -> device forms --------------------------------------------------
    pending = []   # exception raised inside a callback: ctypes cannot unwind through the C frames, so it is kept,
                   # the engine is asked to stop (pc_request_abort) and the exception is re-raised after the call
    keep = []      # ctypes callback objects must outlive the run
    L.pc_request_abort.restype = None

    def _guard(fn, fallback):
        def wrapped(*a):
            if pending:
                return fallback
            try:
                return fn(*a)
            except BaseException as ex:  # noqa: BLE001 -- re-raised below
                pending.append(ex)
                L.pc_request_abort()
                return fallback
        return wrapped

    if isinstance(loglikelihood, _builtin._DeviceLikelihood):
        like_fn = loglikelihood.register(nDims)
    else:
        # polychord.py:581-587 wrap_loglikelihood: the callable returns logL or (logL, phi); theta is read-only
        def _ll(theta_p, nd, phi_p, nder):
            theta = _view(theta_p, (nd,))
            theta.flags.writeable = False
            res = loglikelihood(theta)
            if isinstance(res, tuple):
                logL, phi = res
                if nder:
                    _view(phi_p, (nder,))[:] = phi
            else:
                logL = res
            return float(logL)
        like_fn = _capi.LL_CB(_guard(_ll, float(kwargs['logzero'])))
        keep.append(like_fn)
    addr = _device_prior(prior, nDims, L)
    if addr is not None:
        prior_fn = C.cast(addr, _capi.PRIOR_CB)
    else:
        # polychord.py:589-590 wrap_prior: theta[:] = prior(cube)
        def _prior(cube_p, theta_p, nd):
            cube = _view(cube_p, (nd,))
            cube.flags.writeable = False
            _view(theta_p, (nd,))[:] = prior(cube)
        prior_fn = _capi.PRIOR_CB(_guard(_prior, None))
        keep.append(prior_fn)

    def _dumper(ndead, nlive, npars, live, dead, logweights, logZ, logZerr):
        lv, dd, lw = _view(live, (nlive, npars)), _view(dead, (ndead, npars)), _view(logweights, (ndead,))
        if nlive == 0:  # the final call: every point is dead (nested_sampling.F90:392)
            on_final_dump(dd, lw, logZ, logZerr)
        user_dumper(lv, dd, lw, logZ, logZerr)

    dcb = _capi.DUMPER_CB(_guard(_dumper, None))
    ngrade = len(kwargs['grade_dims'])
    grade_frac = (C.c_double * ngrade)(*[float(f) for f in kwargs['grade_frac']])
    grade_dims = (C.c_int * ngrade)(*kwargs['grade_dims'])
    nl = sorted(kwargs['nlives'].items())
    loglikes = (C.c_double * max(len(nl), 1))(*[k for k, _ in nl])
    nlives = (C.c_int * max(len(nl), 1))(*[v for _, v in nl])
    comm = C.c_int(0)
    L.polychord_c_interface.restype = None
    L.polychord_c_interface.argtypes = _ARGTYPES
    _call(L, like_fn, prior_fn, dcb, kwargs, nDims, nDerived, ngrade, grade_frac, grade_dims, nl, loglikes,
          nlives, comm)
    if pending:
        raise pending[0]


def _call(L, like_fn, prior_fn, dcb, kwargs, nDims, nDerived, ngrade, grade_frac, grade_dims, nl, loglikes, nlives,
          comm):
    L.polychord_c_interface(
        C.cast(like_fn, C.c_void_p), C.cast(prior_fn, C.c_void_p), C.cast(dcb, C.c_void_p),
        int(kwargs['nlive']), int(kwargs['num_repeats']), int(kwargs['nprior']), int(kwargs['nfail']),
        bool(kwargs['do_clustering']), int(kwargs['feedback']), float(kwargs['precision_criterion']),
        float(kwargs['logzero']), int(kwargs['max_ndead']), float(kwargs['boost_posterior']),
        bool(kwargs['posteriors']), bool(kwargs['equals']), bool(kwargs['cluster_posteriors']),
        bool(kwargs['write_resume']), bool(kwargs['write_paramnames']), bool(kwargs['read_resume']),
        bool(kwargs['write_stats']), bool(kwargs['write_live']), bool(kwargs['write_dead']),
        bool(kwargs['write_prior']), bool(kwargs['maximise']), float(kwargs['compression_factor']),
        bool(kwargs['synchronous']), int(nDims), nDerived, str(kwargs['base_dir']).encode(),
        str(kwargs['file_root']).encode(), ngrade, grade_frac, grade_dims, len(nl), loglikes, nlives,
        int(kwargs['seed']), C.byref(comm))


def run_polychord(loglikelihood, nDims, nDerived, settings, prior=default_prior, dumper=default_dumper):
    """Legacy entry point (polychord.py:16): settings object instead of keywords."""
    kw = {k: getattr(settings, k) for k in (
        'nlive', 'num_repeats', 'nprior', 'nfail', 'do_clustering', 'feedback', 'precision_criterion', 'logzero',
        'max_ndead', 'boost_posterior', 'posteriors', 'equals', 'cluster_posteriors', 'write_resume',
        'write_paramnames', 'read_resume', 'write_stats', 'write_live', 'write_dead', 'write_prior', 'maximise',
        'compression_factor', 'synchronous', 'base_dir', 'file_root', 'grade_dims', 'nlives', 'seed')}
    kw['grade_frac'] = settings.grade_frac
    if getattr(settings, 'cube_samples', None) is not None:   # polychord.py:157-158
        kw['cube_samples'] = settings.cube_samples
    lite = run(loglikelihood, nDims, nDerived=nDerived, prior=prior, dumper=dumper, _legacy_output=True, **kw)
    if settings.write_stats:  # polychord.py:218: PolyChordOutput(base_dir, file_root), parsed from <root>.stats
        out = PolyChordOutput(settings.base_dir, settings.file_root)
        out.samples = lite
        return out
    return lite
