"""ctypes binding of the CPU oracle (oracle/pc_oracle.cpp).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
polychordlite_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB_PATH = ORACLE_DIR / "_build" / "liboracle.so"


class OracleSettings(C.Structure):
    _fields_ = [
        ("nDims", C.c_int), ("nDerived", C.c_int), ("nlive", C.c_int), ("num_repeats", C.c_int),
        ("nprior", C.c_int), ("nfail", C.c_int), ("do_clustering", C.c_int),
        ("precision_criterion", C.c_double), ("logzero", C.c_double), ("max_ndead", C.c_int),
        ("boost_posterior", C.c_double), ("posteriors", C.c_int), ("equals", C.c_int),
        ("cluster_posteriors", C.c_int), ("compression_factor", C.c_double), ("seed", C.c_int),
        ("batch_K", C.c_int),
    ]


class OracleResult(C.Structure):
    _fields_ = [
        ("logZ", C.c_double), ("logZerr", C.c_double), ("logZ_raw", C.c_double), ("logZ2_raw", C.c_double),
        ("ndead", C.c_longlong), ("nlike", C.c_longlong), ("nchains", C.c_longlong),
        ("ngenerations", C.c_longlong), ("nupdates", C.c_longlong), ("nfailures", C.c_longlong),
        ("nslices", C.c_longlong), ("nphantoms_final", C.c_longlong), ("seconds", C.c_double),
        ("ncluster", C.c_longlong), ("nsplits", C.c_longlong),
    ]


LL_CB = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int)
PRIOR_CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)
DUMPER_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                        C.POINTER(C.c_double), C.c_double, C.c_double)

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR)], check=True)


def lib():
    global _lib
    if _lib is None:
        src = ORACLE_DIR / "pc_oracle.cpp"
        if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
            build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.oracle_uniform.restype = C.c_double
        _lib.oracle_uniform.argtypes = [C.c_uint, C.c_uint, C.c_ulonglong, C.c_uint, C.c_uint]
        _lib.oracle_inv_normal_cdf.restype = C.c_double
        _lib.oracle_inv_normal_cdf.argtypes = [C.c_double]
        _lib.oracle_logaddexp.restype = C.c_double
        _lib.oracle_logaddexp.argtypes = [C.c_double, C.c_double]
        _lib.oracle_logsumexp.restype = C.c_double
    return _lib


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def make_settings(nDims, nDerived=0, nlive=500, num_repeats=None, nprior=-1, nfail=-1, do_clustering=False,
                  precision_criterion=1e-3, logzero=-1e30, max_ndead=-1, boost_posterior=0.0, posteriors=False,
                  equals=False, cluster_posteriors=False, compression_factor=np.exp(-1), seed=0, batch_K=0):
    s = OracleSettings()
    s.nDims, s.nDerived, s.nlive = nDims, nDerived, nlive
    s.num_repeats = 5 * nDims if num_repeats is None else num_repeats
    s.nprior, s.nfail, s.do_clustering = nprior, nfail, int(do_clustering)
    s.precision_criterion, s.logzero, s.max_ndead = precision_criterion, logzero, max_ndead
    s.boost_posterior, s.posteriors, s.equals = boost_posterior, int(posteriors), int(equals)
    s.cluster_posteriors, s.compression_factor = int(cluster_posteriors), compression_factor
    s.seed, s.batch_K = seed, batch_K
    return s


LIKE_KINDS = {"gaussian": 0, "rastrigin": 1, "corr_gaussian": 2, "callback": 3}


def run(settings, like="gaussian", like_params=None, prior_lo=None, prior_hi=None, ll_cb=None, prior_cb=None,
        want_dump=False):
    """Full run.  Returns (OracleResult, dumps) with dumps a list of dicts when want_dump."""
    L = lib()
    lp = None if like_params is None else np.ascontiguousarray(like_params, dtype=np.float64)
    lo = None if prior_lo is None else np.ascontiguousarray(prior_lo, dtype=np.float64)
    hi = None if prior_hi is None else np.ascontiguousarray(prior_hi, dtype=np.float64)
    dumps = []

    def _dump(ndead, nlive, npars, live, dead, lw, logZ, logZerr):
        dumps.append(dict(
            live=np.ctypeslib.as_array(live, shape=(max(nlive, 1), npars))[:nlive].copy(),
            dead=np.ctypeslib.as_array(dead, shape=(max(ndead, 1), npars))[:ndead].copy(),
            logweights=np.ctypeslib.as_array(lw, shape=(max(ndead, 1),))[:ndead].copy(),
            logZ=logZ, logZerr=logZerr))

    dcb = DUMPER_CB(_dump) if want_dump else C.cast(None, DUMPER_CB)
    res = OracleResult()
    L.oracle_run(C.byref(settings), LIKE_KINDS[like], _dptr(lp), 0 if lp is None else lp.size, _dptr(lo), _dptr(hi),
                 ll_cb if ll_cb is not None else C.cast(None, LL_CB),
                 prior_cb if prior_cb is not None else C.cast(None, PRIOR_CB), dcb, C.byref(res))
    return res, dumps


def set_nlives(schedule=None):
    """Dynamic nlive of the following reference-schedule runs: {loglike threshold: nlive}; None or {} clears."""
    L = lib()
    L.oracle_set_nlives.restype = None
    L.oracle_set_nlives.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    items = sorted((schedule or {}).items())
    ll = (C.c_double * max(len(items), 1))(*[float(k) for k, _ in items])
    nl = (C.c_int * max(len(items), 1))(*[int(v) for _, v in items])
    L.oracle_set_nlives(ll, nl, len(items))


def set_initial_cubes(cubes):
    """cube_samples: the next run() starts from these live points."""
    cubes = np.ascontiguousarray(cubes, dtype=np.float64)
    L = lib()
    L.oracle_set_initial_cubes.restype = None
    L.oracle_set_initial_cubes.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    L.oracle_set_initial_cubes(_dptr(cubes), cubes.shape[0], cubes.shape[1])


def last_clusters():
    """Evidence clusters of the last batched run with clustering: (nactive, array[ncl, 3] of logZp, logZp2, logXp at the
    end of sampling); the clusters alive at the end come first, then the deleted ones in order of deletion."""
    L = lib()
    L.oracle_last_clusters.restype = C.c_int
    L.oracle_last_clusters.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double)]
    nact = C.c_int(0)
    n = L.oracle_last_clusters(C.byref(nact), None)
    out = np.zeros((max(n, 1), 3))
    L.oracle_last_clusters(C.byref(nact), _dptr(out))
    return int(nact.value), out[:n]


def last_dead_clusters():
    """(identity of the cluster each dead point of the last clustered batched run died in, parent of every identity,
    identity of every cluster last_clusters() lists)."""
    L = lib()
    L.oracle_last_dead_clusters.restype = C.c_longlong
    L.oracle_last_dead_clusters.argtypes = [C.POINTER(C.c_int), C.c_longlong]
    n = L.oracle_last_dead_clusters(None, 0)
    out = np.zeros(max(n, 1), dtype=np.int32)
    L.oracle_last_dead_clusters(out.ctypes.data_as(C.POINTER(C.c_int)), n)
    L.oracle_last_cluster_tree.restype = C.c_int
    L.oracle_last_cluster_tree.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    m = L.oracle_last_cluster_tree(None, None)
    par = np.zeros(max(m, 1), dtype=np.int32)
    ncl = L.oracle_last_clusters(None, None)
    cu = np.zeros(max(ncl, 1), dtype=np.int32)
    L.oracle_last_cluster_tree(par.ctypes.data_as(C.POINTER(C.c_int)), cu.ctypes.data_as(C.POINTER(C.c_int)))
    return out[:n], par[:m], cu[:ncl]


def last_boosted(npars):
    """Phantoms the last run() promoted to posterior samples: (rows[nb, npars], dead_index[nb], logw[nb])."""
    L = lib()
    L.oracle_last_boosted.restype = C.c_longlong
    L.oracle_last_boosted.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                                      C.c_longlong, C.c_int]
    nb = L.oracle_last_boosted(None, None, None, 0, npars)
    rows, idx, lw = np.zeros((max(nb, 1), npars)), np.zeros(max(nb, 1), dtype=np.int64), np.zeros(max(nb, 1))
    L.oracle_last_boosted(_dptr(rows), idx.ctypes.data_as(C.POINTER(C.c_longlong)), _dptr(lw), nb, npars)
    return rows[:nb], idx[:nb], lw[:nb]


def slice_chain(settings, seed_point, cholesky, logL, uid, like="gaussian", like_params=None, prior_lo=None,
                prior_hi=None):
    L = lib()
    D, P, R = settings.nDims, settings.nDerived, settings.num_repeats
    T = 2 * D + P + 2
    lp = None if like_params is None else np.ascontiguousarray(like_params, dtype=np.float64)
    lo = None if prior_lo is None else np.ascontiguousarray(prior_lo, dtype=np.float64)
    hi = None if prior_hi is None else np.ascontiguousarray(prior_hi, dtype=np.float64)
    sp = np.ascontiguousarray(seed_point, dtype=np.float64)
    ch = np.asfortranarray(cholesky, dtype=np.float64)
    babies = np.zeros((R, T))
    n = C.c_longlong(0)
    L.oracle_slice_chain(C.byref(settings), LIKE_KINDS[like], _dptr(lp), 0 if lp is None else lp.size, _dptr(lo),
                         _dptr(hi), _dptr(sp), ch.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(logL),
                         C.c_ulonglong(uid), _dptr(babies), C.byref(n))
    return babies, n.value


def generate_nhats(settings, uid):
    L = lib()
    out = np.zeros((settings.num_repeats, settings.nDims))
    L.oracle_generate_nhats(C.byref(settings), C.c_ulonglong(uid), _dptr(out))
    return out  # row i = direction i


def calculate_points(settings, cubes, like="gaussian", like_params=None, prior_lo=None, prior_hi=None):
    L = lib()
    D, P = settings.nDims, settings.nDerived
    T = 2 * D + P + 2
    cubes = np.atleast_2d(np.asarray(cubes, dtype=np.float64))
    rec = np.zeros((cubes.shape[0], T))
    rec[:, :D] = cubes
    lp = None if like_params is None else np.ascontiguousarray(like_params, dtype=np.float64)
    lo = None if prior_lo is None else np.ascontiguousarray(prior_lo, dtype=np.float64)
    hi = None if prior_hi is None else np.ascontiguousarray(prior_hi, dtype=np.float64)
    n = L.oracle_calculate_points(C.byref(settings), LIKE_KINDS[like], _dptr(lp), 0 if lp is None else lp.size,
                                  _dptr(lo), _dptr(hi), _dptr(rec), rec.shape[0])
    return rec, n


def philox(ctr, key):
    L = lib()
    c = (C.c_uint * 4)(*ctr)
    k = (C.c_uint * 2)(*key)
    o = (C.c_uint * 4)()
    L.oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def calc_cholesky(a):
    L = lib()
    a = np.asfortranarray(a, dtype=np.float64)
    out = np.zeros_like(a, order="F")
    L.oracle_calc_cholesky(a.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)),
                           a.shape[0])
    return np.array(out)


def evidence_sequence(state, logLs, n0, dec, logzero=-1e30):
    L = lib()
    st = np.array(state, dtype=np.float64)
    ll = np.ascontiguousarray(logLs, dtype=np.float64)
    lw = np.zeros(ll.size)
    L.oracle_evidence_sequence(_dptr(st), _dptr(ll), ll.size, n0, dec, C.c_double(logzero), _dptr(lw))
    return st, lw


def random_inverse_covmat(seed, D, sigma):
    L = lib()
    inv = np.zeros((D, D), order="F")
    ld = C.c_double(0)
    L.oracle_random_inverse_covmat(C.c_uint(seed), D, C.c_double(sigma), inv.ctypes.data_as(C.POINTER(C.c_double)),
                                   C.byref(ld))
    return np.array(inv), ld.value


def nn_clustering(points):
    """NN_clustering (clustering.f90:15-97) of the rows of `points`; returns (labels, number of clusters)."""
    L = lib()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    m, D = pts.shape
    labels = np.zeros(m, dtype=np.int32)
    L.oracle_nn_clustering.restype = C.c_int
    num = L.oracle_nn_clustering(_dptr(pts), m, D, labels.ctypes.data_as(C.POINTER(C.c_int)))
    return labels, num


def set_grades(grade_dims=(), grade_repeats=()):
    """oracle_set_grades: fast/slow grades for the following runs / chain probes; no arguments clears them."""
    n = len(grade_dims)
    lib().oracle_set_grades(n, (C.c_int * max(n, 1))(*grade_dims), (C.c_int * max(n, 1))(*grade_repeats))
