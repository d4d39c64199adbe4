"""ctypes binding of libchord.so -- the C ABI declared in include/polychord_b200.h.

Nothing here computes: every call goes into the CUDA library.  If the library is missing the
import fails loudly (there is no CPU fallback by design).
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("PC_LIBCHORD", PKG / "lib" / "libchord.so"))

LL_CB = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int)
PRIOR_CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)
DUMPER_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                        C.POINTER(C.c_double), C.c_double, C.c_double)

LIKE_KINDS = {"gaussian": 0, "rastrigin": 1, "corr_gaussian": 2}

# every symbol include/polychord_b200.h declares
EXPORTS = [
    "polychord_c_interface", "polychord_c_interface_ini",
    "pc_register_device_likelihood", "pc_register_device_prior", "pc_clear_registrations",
    "pc_gaussian_loglikelihood", "pc_rastrigin_loglikelihood", "pc_corr_gaussian_loglikelihood",
    "pc_unit_prior", "pc_uniform_prior", "pc_set_option", "pc_get_option", "pc_set_stream", "pc_release_memory", "pc_mgpu_create", "pc_mgpu_attach", "pc_mgpu_destroy",
    "pc_last_run_info", "pc_run", "pc_run_ensemble", "pc_slice_chains", "pc_calculate_points",
    "pc_device_philox", "pc_device_uniforms", "pc_device_inv_normal_cdf", "pc_device_directions",
    "pc_device_evidence", "pc_device_cholesky", "pc_device_count", "pc_version", "pc_format_e24", "pc_write_files", "pc_request_abort", "pc_cluster_points", "pc_set_grades", "pc_set_resume", "pc_measure_fp64_tflops", "pc_ini_prior_transform", "pc_last_boosted", "pc_maximise", "pc_prior_log_density", "pc_set_initial_live", "pc_write_files_boosted", "pc_set_nlives", "pc_resume_text_probe", "pc_last_clusters", "pc_last_dead_clusters", "pc_last_cluster_tree", "pc_auto_batch_size",
]


class Settings(C.Structure):
    _fields_ = [
        ("nDims", C.c_int), ("nDerived", C.c_int), ("nlive", C.c_int), ("num_repeats", C.c_int),
        ("nprior", C.c_int), ("nfail", C.c_int), ("do_clustering", C.c_int), ("feedback", C.c_int),
        ("precision_criterion", C.c_double), ("logzero", C.c_double), ("max_ndead", C.c_int),
        ("boost_posterior", C.c_double), ("posteriors", C.c_int), ("equals", C.c_int),
        ("cluster_posteriors", C.c_int), ("compression_factor", C.c_double), ("seed", C.c_int),
    ]


class RunInfo(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("logZ", C.c_double), ("logZerr", C.c_double), ("logZ_raw", C.c_double),
        ("logZ2_raw", C.c_double), ("ndead", C.c_longlong), ("nlike", C.c_longlong), ("nchains", C.c_longlong),
        ("ngenerations", C.c_longlong), ("nupdates", C.c_longlong), ("nfailures", C.c_longlong),
        ("nslices", C.c_longlong), ("nphantoms_final", C.c_longlong), ("batch_K", C.c_int),
        ("warps_per_cta", C.c_int), ("ctas_per_run", C.c_int), ("kernel_launches", C.c_int),
        ("device_ms", C.c_double), ("wall_ms", C.c_double), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
        ("algorithmic_bytes", C.c_longlong), ("phase_ms", C.c_double * 8),
        ("ncluster_max", C.c_longlong), ("ncluster_updates", C.c_longlong), ("cluster_ms", C.c_double),
        ("kernel_G", C.c_int), ("kernel_DPL", C.c_int), ("kernel_kind", C.c_int), ("kernel_mode", C.c_int),
        ("nlive_final", C.c_int), ("pad_", C.c_int),
    ]

    @property
    def kernel(self):
        """Name of the run kernel that was dispatched (as ncu lists it)."""
        return f"pc_run_kernel<{self.kernel_G},{self.kernel_DPL},{self.kernel_kind},{self.kernel_mode}>"

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["phase_ms"] = dict(zip(("wait_chains", "phase_S", "finish_update", "phase_U", "prep", "whiten", "slices",
                                  "kernel"), list(self.phase_ms)))
        return d


_lib = None


def lib():
    """Load libchord.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m polychordlite_b200._build` "
                              "(the engine has no CPU fallback)")
        L = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
        L.pc_version.restype = C.c_char_p
        L.pc_get_option.restype = C.c_double
        L.pc_set_option.argtypes = [C.c_char_p, C.c_double]
        L.pc_get_option.argtypes = [C.c_char_p]
        L.pc_set_stream.argtypes = [C.c_void_p]
        for name in ("pc_gaussian_loglikelihood", "pc_rastrigin_loglikelihood", "pc_corr_gaussian_loglikelihood"):
            getattr(L, name).restype = C.c_double
        _lib = L
    return _lib


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _arr(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ravel()


def make_settings(nDims, nDerived=0, nlive=500, num_repeats=None, nprior=-1, nfail=-1, do_clustering=False, feedback=0,
                  precision_criterion=1e-3, logzero=-1e30, max_ndead=-1, boost_posterior=0.0, posteriors=False,
                  equals=False, cluster_posteriors=False, compression_factor=np.exp(-1), seed=0):
    s = Settings()
    s.nDims, s.nDerived, s.nlive = nDims, nDerived, nlive
    s.num_repeats = 5 * nDims if num_repeats is None else num_repeats
    s.nprior, s.nfail, s.do_clustering, s.feedback = nprior, nfail, int(do_clustering), feedback
    s.precision_criterion, s.logzero, s.max_ndead = precision_criterion, logzero, max_ndead
    s.boost_posterior, s.posteriors, s.equals = boost_posterior, int(posteriors), int(equals)
    s.cluster_posteriors, s.compression_factor, s.seed = int(cluster_posteriors), compression_factor, seed
    return s


def set_option(name, value):
    if lib().pc_set_option(name.encode(), float(value)) != 0:
        raise KeyError(name)


def get_option(name):
    return lib().pc_get_option(name.encode())


def device_count():
    return lib().pc_device_count()


def _prior_params(lo, hi):
    if lo is None or hi is None:
        return None
    return np.concatenate([_arr(lo), _arr(hi)])


def run(settings, like="gaussian", like_params=None, prior_lo=None, prior_hi=None, want_dump=False, abort_after_dumps=None):
    """One device run (pc_run).  Returns (RunInfo, dumps).  abort_after_dumps: call pc_request_abort() from inside the
    N-th dumper call (an interrupted run)."""
    L = lib()
    lp, pp = _arr(like_params), _prior_params(prior_lo, prior_hi)
    dumps = []

    def _dump(ndead, nlive, npars, live, dead, lw, logZ, logZerr):
        if abort_after_dumps is not None and len(dumps) + 1 == abort_after_dumps:
            L.pc_request_abort()
        dumps.append(dict(
            live=np.ctypeslib.as_array(live, shape=(max(nlive, 1), npars))[:nlive].copy(),
            dead=np.ctypeslib.as_array(dead, shape=(max(ndead, 1), npars))[:ndead].copy(),
            logweights=np.ctypeslib.as_array(lw, shape=(max(ndead, 1),))[:ndead].copy(),
            logZ=logZ, logZerr=logZerr))

    dcb = DUMPER_CB(_dump) if (want_dump or abort_after_dumps) else C.cast(None, DUMPER_CB)
    info = RunInfo()
    rc = L.pc_run(C.byref(settings), LIKE_KINDS[like], _dptr(lp), 0 if lp is None else lp.size, _dptr(pp),
                  0 if pp is None else pp.size, dcb, C.byref(info))
    if rc != 0:
        raise RuntimeError(f"pc_run failed with status {rc}")
    return info, dumps


def set_initial_live(cube_samples=None):
    """pc_set_initial_live: the next run through polychord_c_interface starts from these cube points (None clears)."""
    L = lib()
    L.pc_set_initial_live.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    if cube_samples is None:
        L.pc_set_initial_live(None, 0, 0)
        return
    a = np.ascontiguousarray(cube_samples, dtype=np.float64)
    if a.ndim != 2 or L.pc_set_initial_live(_dptr(a), a.shape[0], a.shape[1]) != 0:
        raise ValueError("cube_samples must be a (npoints, nDims) array")


def set_nlives(schedule=None):
    """pc_set_nlives: dynamic nlive of the following pc_run() calls, {loglike threshold: nlive}; None or {} clears."""
    L = lib()
    L.pc_set_nlives.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    items = sorted((schedule or {}).items())
    ll = (C.c_double * max(len(items), 1))(*[float(k) for k, _ in items])
    nl = (C.c_int * max(len(items), 1))(*[int(v) for _, v in items])
    if L.pc_set_nlives(ll, nl, len(items)) != 0:
        raise ValueError("the nlives schedule holds at most 16 entries")


def last_boosted(npars):
    """Phantoms the last run promoted to posterior samples (boost_posterior): (rows[nb, npars], dead_index[nb], logw[nb])."""
    L = lib()
    L.pc_last_boosted.restype = C.c_longlong
    L.pc_last_boosted.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.c_longlong, C.c_int]
    nb = L.pc_last_boosted(None, None, None, 0, npars)
    if nb < 0:
        raise ValueError("npars does not match the last run")
    rows, idx, lw = np.zeros((max(nb, 1), npars)), np.zeros(max(nb, 1), dtype=np.int64), np.zeros(max(nb, 1))
    L.pc_last_boosted(_dptr(rows), idx.ctypes.data_as(C.POINTER(C.c_longlong)), _dptr(lw), nb, npars)
    return rows[:nb], idx[:nb], lw[:nb]


def run_ensemble(settings, seeds, like="gaussian", like_params=None, prior_lo=None, prior_hi=None):
    L = lib()
    lp, pp = _arr(like_params), _prior_params(prior_lo, prior_hi)
    seeds = np.ascontiguousarray(seeds, dtype=np.int32)
    infos = (RunInfo * len(seeds))()
    rc = L.pc_run_ensemble(C.byref(settings), LIKE_KINDS[like], _dptr(lp), 0 if lp is None else lp.size, _dptr(pp),
                           0 if pp is None else pp.size, len(seeds), seeds.ctypes.data_as(C.POINTER(C.c_int)), infos)
    if rc != 0:
        raise RuntimeError(f"pc_run_ensemble failed with status {rc}")
    return list(infos)


def slice_chains(settings, seed_points, cholesky, logL, uid, like="gaussian", like_params=None, prior_lo=None,
                 prior_hi=None):
    L = lib()
    D, P, R = settings.nDims, settings.nDerived, settings.num_repeats
    T = 2 * D + P + 2
    sp = np.ascontiguousarray(seed_points, dtype=np.float64).reshape(-1, T)
    nch = sp.shape[0]
    ch = np.asfortranarray(cholesky, dtype=np.float64)
    ll = np.ascontiguousarray(np.broadcast_to(np.asarray(logL, dtype=np.float64), (nch,)))
    ui = np.ascontiguousarray(uid, dtype=np.uint64)
    lp, pp = _arr(like_params), _prior_params(prior_lo, prior_hi)
    babies = np.zeros((nch, R, T))
    nlike = np.zeros(nch, dtype=np.int64)
    rc = L.pc_slice_chains(C.byref(settings), LIKE_KINDS[like], _dptr(lp), 0 if lp is None else lp.size, _dptr(pp),
                           0 if pp is None else pp.size, nch, _dptr(sp), ch.ctypes.data_as(C.POINTER(C.c_double)),
                           _dptr(ll), ui.ctypes.data_as(C.POINTER(C.c_ulonglong)), _dptr(babies),
                           nlike.ctypes.data_as(C.POINTER(C.c_longlong)))
    if rc != 0:
        raise RuntimeError(f"pc_slice_chains failed with status {rc}")
    return babies, nlike


def calculate_points(settings, cubes, like="gaussian", like_params=None, prior_lo=None, prior_hi=None):
    L = lib()
    D, P = settings.nDims, settings.nDerived
    T = 2 * D + P + 2
    cubes = np.atleast_2d(np.asarray(cubes, dtype=np.float64))
    rec = np.zeros((cubes.shape[0], T))
    rec[:, :D] = cubes
    lp, pp = _arr(like_params), _prior_params(prior_lo, prior_hi)
    n = L.pc_calculate_points(C.byref(settings), LIKE_KINDS[like], _dptr(lp), 0 if lp is None else lp.size, _dptr(pp),
                              0 if pp is None else pp.size, _dptr(rec), rec.shape[0])
    if n < 0:
        raise RuntimeError(f"pc_calculate_points failed with status {n}")
    return rec, n


def device_philox(ctr, key):
    c = (C.c_uint * 4)(*ctr)
    k = (C.c_uint * 2)(*key)
    o = (C.c_uint * 4)()
    if lib().pc_device_philox(c, k, o) != 0:
        raise RuntimeError("pc_device_philox failed")
    return [int(x) for x in o]


def device_uniforms(seed, tag, uid, a0, b, n):
    out = np.zeros(n)
    if lib().pc_device_uniforms(C.c_uint(seed), C.c_uint(tag), C.c_ulonglong(uid), C.c_uint(a0), C.c_uint(b), n,
                                _dptr(out)) != 0:
        raise RuntimeError("pc_device_uniforms failed")
    return out


def device_inv_normal_cdf(p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.zeros_like(p)
    if lib().pc_device_inv_normal_cdf(_dptr(p), p.size, _dptr(out)) != 0:
        raise RuntimeError("pc_device_inv_normal_cdf failed")
    return out


def device_directions(nDims, num_repeats, seed, uid):
    out = np.zeros((num_repeats, nDims))
    if lib().pc_device_directions(nDims, num_repeats, C.c_uint(seed), C.c_ulonglong(uid), _dptr(out)) != 0:
        raise RuntimeError("pc_device_directions failed")
    return out


def device_evidence(state, logLs, n_start):
    st = np.array(state, dtype=np.float64)
    ll = np.ascontiguousarray(logLs, dtype=np.float64)
    lw = np.zeros(ll.size)
    if lib().pc_device_evidence(_dptr(st), _dptr(ll), ll.size, n_start, _dptr(lw)) != 0:
        raise RuntimeError("pc_device_evidence failed")
    return st, lw


def device_cholesky(a):
    a = np.asfortranarray(a, dtype=np.float64)
    out = np.zeros_like(a, order="F")
    fb = lib().pc_device_cholesky(a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0],
                                  out.ctypes.data_as(C.POINTER(C.c_double)))
    if fb < 0:
        raise RuntimeError("pc_device_cholesky failed")
    return np.array(out), fb


def mgpu_create(settings, world):
    """This rank's exchange block for a sharded run; returns its 64-byte CUDA IPC handle."""
    h = (C.c_ubyte * 64)()
    if lib().pc_mgpu_create(C.byref(settings), int(world), h) != 0:
        raise RuntimeError("pc_mgpu_create failed")
    return bytes(h)


def mgpu_attach(rank, world, handles):
    buf = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
    if lib().pc_mgpu_attach(int(rank), int(world), buf) != 0:
        raise RuntimeError("pc_mgpu_attach failed")


def mgpu_destroy():
    lib().pc_mgpu_destroy()


def last_run_info():
    info = RunInfo()
    lib().pc_last_run_info(C.byref(info))
    return info


# ---- output files (host only) ----------------------------------------------------------------------
FILE_FLAGS = {"stats": 1, "live": 2, "dead": 4, "prior": 8, "posteriors": 16, "equals": 32}


def format_e24(value):
    """Fortran E24.15E3 rendering of a double (utils.F90:19), as the engine writes every number."""
    buf = C.create_string_buffer(25)
    L = lib()
    L.pc_format_e24.argtypes = [C.c_double, C.c_char_p]
    L.pc_format_e24.restype = None
    L.pc_format_e24(float(value), buf)
    return buf.value.decode()


def last_clusters():
    """Clusters of the last run with do_clustering: (nactive, rows[ncl, 2] of log<Z_p>, log<Z_p^2>, uid[ncl])."""
    L = lib()
    L.pc_last_clusters.restype = C.c_int
    L.pc_last_clusters.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    nact = C.c_int(0)
    n = L.pc_last_clusters(C.byref(nact), None, None)
    rows = np.zeros((max(n, 1), 2))
    uid = np.zeros(max(n, 1), dtype=np.int32)
    L.pc_last_clusters(C.byref(nact), _dptr(rows), uid.ctypes.data_as(C.POINTER(C.c_int)))
    return int(nact.value), rows[:n], uid[:n]


def last_dead_clusters():
    """(identity of the cluster each dead point of the last clustered run died in, parent of every identity)."""
    L = lib()
    L.pc_last_dead_clusters.restype = C.c_longlong
    L.pc_last_dead_clusters.argtypes = [C.POINTER(C.c_int), C.c_longlong]
    n = L.pc_last_dead_clusters(None, 0)
    out = np.zeros(max(n, 1), dtype=np.int32)
    L.pc_last_dead_clusters(out.ctypes.data_as(C.POINTER(C.c_int)), n)
    L.pc_last_cluster_tree.restype = C.c_int
    L.pc_last_cluster_tree.argtypes = [C.POINTER(C.c_int)]
    m = L.pc_last_cluster_tree(None)
    par = np.zeros(max(m, 1), dtype=np.int32)
    L.pc_last_cluster_tree(par.ctypes.data_as(C.POINTER(C.c_int)))
    return out[:n], par[:m]


def resume_text_probe(path, out_path=None):
    """pc_resume_text_probe: parse a resume file in the reference's text layout (host only); optionally re-write it."""
    L = lib()
    L.pc_resume_text_probe.argtypes = [C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.c_char_p]
    L.pc_resume_text_probe.restype = C.c_int
    ints = (C.c_longlong * 8)()
    reals = (C.c_double * 6)()
    rc = L.pc_resume_text_probe(str(path).encode(), ints, reals, str(out_path).encode() if out_path else None)
    if rc != 0:
        raise ValueError(f"pc_resume_text_probe: status {rc} (malformed resume file; the message is on stderr)")
    keys_i = ["nDims", "nDerived", "ndead", "ncluster", "ncluster_dead", "nlive", "nphantom", "nlike"]
    keys_r = ["logZ", "logZ2", "logX", "logX_last_update", "logL_min", "logL_max"]
    return {**{k: int(v) for k, v in zip(keys_i, ints)}, **{k: float(v) for k, v in zip(keys_r, reals)}}


def write_files(base_dir, file_root, nDims, nDerived, dead_rows, dead_logw, live_rows, logZ, logZerr, nlike,
                num_repeats, compression_factor=float(np.exp(-1)), seed=0, flags=("stats", "live", "dead", "posteriors",
                                                                                  "equals"), boosted=None):
    """pc_write_files: the engine's file writer driven with explicit arrays (no device needed).
    boosted = (rows, logw, after): pc_write_files_boosted."""
    L = lib()
    if boosted is not None:
        brows = np.ascontiguousarray(boosted[0], dtype=np.float64).reshape(-1, nDims + nDerived + 2)
        blogw = np.ascontiguousarray(boosted[1], dtype=np.float64)
        bafter = np.ascontiguousarray(boosted[2], dtype=np.int64)
        dead_rows = np.ascontiguousarray(dead_rows, dtype=np.float64).reshape(-1, nDims + nDerived + 2)
        live_rows = np.ascontiguousarray(live_rows, dtype=np.float64).reshape(-1, nDims + nDerived + 2)
        dead_logw = np.ascontiguousarray(dead_logw, dtype=np.float64)
        L.pc_write_files_boosted.restype = C.c_int
        L.pc_write_files_boosted.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                             C.c_double, C.c_double, C.c_longlong, C.c_int, C.c_double, C.c_uint,
                                             C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_longlong)]
        fl = sum(FILE_FLAGS[f] for f in flags)
        return L.pc_write_files_boosted(str(base_dir).encode(), str(file_root).encode(), fl, nDims, nDerived,
                                        dead_rows.shape[0], _dptr(dead_rows), _dptr(dead_logw), live_rows.shape[0],
                                        _dptr(live_rows), float(logZ), float(logZerr), int(nlike), int(num_repeats),
                                        float(compression_factor), int(seed), brows.shape[0], _dptr(brows), _dptr(blogw),
                                        bafter.ctypes.data_as(C.POINTER(C.c_longlong)))
    dead_rows = np.ascontiguousarray(dead_rows, dtype=np.float64).reshape(-1, nDims + nDerived + 2)
    live_rows = np.ascontiguousarray(live_rows, dtype=np.float64).reshape(-1, nDims + nDerived + 2)
    dead_logw = np.ascontiguousarray(dead_logw, dtype=np.float64)
    L.pc_write_files.restype = C.c_int
    L.pc_write_files.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double,
                                 C.c_longlong, C.c_int, C.c_double, C.c_uint]
    fl = sum(FILE_FLAGS[f] for f in flags)
    return L.pc_write_files(str(base_dir).encode(), str(file_root).encode(), fl, nDims, nDerived, dead_rows.shape[0],
                            _dptr(dead_rows), _dptr(dead_logw), live_rows.shape[0], _dptr(live_rows), float(logZ),
                            float(logZerr), int(nlike), int(num_repeats), float(compression_factor), int(seed))


def cluster_points(points):
    """pc_cluster_points: (labels, number of clusters) of the rows of `points`."""
    L = lib()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    m, D = pts.shape
    labels = np.zeros(m, dtype=np.int32)
    L.pc_cluster_points.restype = C.c_int
    num = L.pc_cluster_points(_dptr(pts), m, D, labels.ctypes.data_as(C.POINTER(C.c_int)))
    if num < 0:
        raise RuntimeError(f"pc_cluster_points failed with status {num}")
    return labels, num


def set_grades(grade_dims=(), grade_repeats=()):
    """pc_set_grades: fast/slow grades for the following runs / chain probes; no arguments clears them."""
    n = len(grade_dims)
    d = (C.c_int * max(n, 1))(*grade_dims)
    r = (C.c_int * max(n, 1))(*grade_repeats)
    if lib().pc_set_grades(n, d, r) != 0:
        raise ValueError("pc_set_grades failed")


def set_resume(path=None, write=False, read=False):
    """pc_set_resume: resume file of the following runs (path ending in .resume); no arguments clears it."""
    L = lib()
    L.pc_set_resume.argtypes = [C.c_char_p, C.c_int, C.c_int]
    if L.pc_set_resume(None if path is None else str(path).encode(), int(write), int(read)) != 0:
        raise ValueError("pc_set_resume: the path must end in .resume")


def auto_batch_size(nlive, world=1):
    """pc_auto_batch_size: the deaths per generation the engine picks for a run alone on the device / sharded over `world`."""
    L = lib()
    L.pc_auto_batch_size.restype = C.c_int
    L.pc_auto_batch_size.argtypes = [C.c_int, C.c_int]
    return int(L.pc_auto_batch_size(int(nlive), int(world)))


def measure_fp64_tflops():
    L = lib()
    L.pc_measure_fp64_tflops.restype = C.c_double
    return float(L.pc_measure_fp64_tflops())
