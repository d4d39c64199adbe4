"""CPU restatement (numpy/scipy, test infrastructure only) of the reference's prior transforms,
src/polychord/priors.f90, as the .ini driver applies them block by block (hypercube_to_physical, :494-556).

Only tests/ may import this file; nothing under polychordlite_b200/ does.  Parity unpinned by the reference (it
holds no golden vectors for its priors): pinned here by the closed forms themselves (scipy's inverse normal CDF
against AS241) and by the distributional properties tests/test_ini_priors.py checks.

A block is (prior_type, [params of parameter 1, params of parameter 2, ...]) exactly as add_param_to_prior
(:751-789) concatenates them; `cube` is the block's hypercube coordinates.
"""
import numpy as np
from scipy.special import ndtri


def _pairs(q):
    q = np.asarray(q, float)
    return q[0::2], q[1::2]


def uniform_htp(u, q):            # priors.f90:40-55
    a, b = _pairs(q)
    return a + (b - a) * u


def gaussian_htp(u, q):           # priors.f90:73-88
    mu, sig = _pairs(q)
    return mu + sig * ndtri(u)


def log_uniform_htp(u, q):        # priors.f90:114-128
    a, b = _pairs(q)
    return a * (b / a) ** u


def power_uniform_htp(u, q):      # priors.f90:151-170
    q = np.asarray(q, float)
    lo, hi, p = q[0::3], q[1::3], q[2::3]
    a, b = lo ** (1 / p), hi ** (1 / p)
    return (a - u * np.abs(a - b)) ** p


def half_gaussian_htp(u, q):      # priors.f90:172-190
    mu, sig = _pairs(q)
    return mu + sig * ndtri(0.5 + 0.5 * u)


def exponential_htp(u, q):        # priors.f90:192-207
    return -np.log(1 - u) / np.asarray(q, float)


def sort_hypercube(u):            # priors.f90:242-264
    u = np.asarray(u, float)
    n = len(u)
    s = np.empty(n)
    if n == 0:
        return s
    s[n - 1] = u[n - 1] ** (1.0 / n)
    for k in range(n - 2, -1, -1):
        s[k] = u[k] ** (1.0 / (k + 1)) * s[k + 1]
    return s


def adaptive_sorted_transform(u):  # priors.f90:367-385
    u = np.asarray(u, float)
    t = u.copy()
    t[0] = 0.5 + u[0] * (len(u) - 1)
    nfunc = min(int(t[0] + 0.5), len(u) - 1)   # the reference runs past the block at u[0] == 1 exactly
    t[1:nfunc + 1] = sort_hypercube(t[1:nfunc + 1])
    return t


_BASE = {1: uniform_htp, 4: gaussian_htp, 5: half_gaussian_htp, 6: exponential_htp}


def _adaptive(u, q, base, skip):   # priors.f90:389-461: parameters(3:) (or (2:) for the exponential)
    t = adaptive_sorted_transform(u)
    t[1:] = _BASE[base](t[1:], q[skip:])
    return t


def block_htp(prior_type, u, q):
    u = np.asarray(u, float)
    q = np.asarray(q, float)
    if prior_type in (1, 4, 5, 6):
        return _BASE[prior_type](u, q)
    if prior_type == 2:
        return log_uniform_htp(u, q)
    if prior_type == 3:
        return power_uniform_htp(u, q)
    if 7 <= prior_type <= 10:      # sorted families, priors.f90:266-360
        return _BASE[(1, 4, 5, 6)[prior_type - 7]](sort_hypercube(u), q)
    if 11 <= prior_type <= 14:
        base = (1, 4, 5, 6)[prior_type - 11]
        return _adaptive(u, q, base, 1 if base == 6 else 2)
    if prior_type == 15:           # nn_adaptive_layer_gaussian_htp, priors.f90:469-488
        out = u.copy()
        out[0] = 0.5 + 2 * u[0]
        out[1:] = _adaptive(u[1:], q[2:], 5 if out[0] < 1.5 else 4, 2)
        return out
    raise ValueError(prior_type)
