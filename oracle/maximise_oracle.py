"""CPU restatement (numpy, test infrastructure only) of the reference's maximiser: src/polychord/nelder_mead.f90
(nelder_mead :4-75, det :161-211) and src/polychord/maximiser.F90 (dXdtheta :172-202, do_maximisation :80-153).

Only tests/ may import this file.  Parity unpinned by the reference (no golden vectors for the maximiser): pinned by
the analytic maxima tests/test_maximise.py checks.
"""
import numpy as np


def det(matrix):                                     # nelder_mead.f90:161-211
    a = np.array(matrix, dtype=float)
    n = a.shape[0]
    sign = 1.0
    for k in range(n - 1):
        if a[k, k] == 0:
            for i in range(k + 1, n):
                if a[i, k] != 0:
                    a[[i, k], :] = a[[k, i], :]
                    sign = -sign
                    break
            else:
                return 0.0
        for j in range(k + 1, n):
            m = a[j, k] / a[k, k]
            a[j, k + 1:] -= m * a[k, k + 1:]
    return sign * float(np.prod(np.diag(a)))


def nelder_mead(func, x, f, dl, max_iter=200000):    # nelder_mead.f90:4-75; x: (n, n+1) vertices in columns; maximises
    x = np.array(x, dtype=float)
    f = np.array(f, dtype=float)
    n = len(f) - 1
    expo = float(np.float32(1.0) / np.float32(n))    # (det1/det0)**(1./n): a single-precision exponent
    det0 = -1.0
    ncall = 0
    for _ in range(max_iter):
        i = np.argsort(f, kind="stable")
        det1 = abs(det(x[:, i[:n]] - x[:, [i[n]]]))
        if det0 < 0:
            det0 = det1
        with np.errstate(all="ignore"):
            if f[i[n]] - f[i[0]] < dl or (det1 / det0) ** expo < dl:
                break
        xo = x[:, i[1:]].sum(axis=1) / n
        xr = xo + 1.0 * (xo - x[:, i[0]])
        fr = func(xr); ncall += 1
        if fr <= f[i[n]] and f[i[1]] < fr:
            f[i[0]], x[:, i[0]] = fr, xr
        elif fr > f[i[n]]:
            xe = xo + 2.0 * (xr - xo)
            fe = func(xe); ncall += 1
            if fe > fr:
                f[i[0]], x[:, i[0]] = fe, xe
            else:
                f[i[0]], x[:, i[0]] = fr, xr
        else:
            xc = xo + 0.5 * (x[:, i[0]] - xo)
            fc = func(xc); ncall += 1
            if fc > f[i[0]]:
                f[i[0]], x[:, i[0]] = fc, xc
            else:
                for j in range(n):
                    x[:, i[j]] = x[:, i[n]] + 0.5 * (x[:, i[j]] - x[:, i[n]])
                    f[i[j]] = func(x[:, i[j]]); ncall += 1
    i = np.argsort(f, kind="stable")
    return x[:, i[n]].copy(), ncall


def dXdtheta(prior, cube, dx=1e-5):                  # maximiser.F90:172-202
    cube = np.asarray(cube, dtype=float)
    n = len(cube)
    th0 = np.asarray(prior(cube), dtype=float)
    dtheta = np.zeros((n, n))
    s = 1
    for i in range(n):
        c0 = cube.copy()
        if c0[i] + dx >= 1:
            c0[i] -= dx
            s = -s
        else:
            c0[i] += dx
        dtheta[:, i] = np.asarray(prior(c0), dtype=float) - th0
    return n * np.log(dx) - np.log(s * det(dtheta))


def do_maximisation(loglike, prior, live_cube, live_logL, logzero, posterior):   # maximiser.F90:80-153, one cluster
    live_cube = np.asarray(live_cube, dtype=float)
    D = live_cube.shape[1]
    l = np.array(live_logL, dtype=float)
    if posterior:
        l = l + np.array([dXdtheta(prior, c) for c in live_cube])
    i = np.argsort(l, kind="stable")
    top = i[len(l) - D - 1:]
    simplex = live_cube[top].T.copy()
    f = l[top].copy()

    def func(x):
        if np.any(x < 0) or np.any(x > 1):
            return logzero
        v = loglike(np.asarray(prior(x), dtype=float))
        if posterior and v > logzero:
            v += dXdtheta(prior, x)
        return v
    x, _ = nelder_mead(func, simplex, f, 1e-5)
    return x
