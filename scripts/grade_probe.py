import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as gpu
import oracle_lib as oracle
for D, dims, reps, like in [(33, [22, 10, 1], [5, 5, 6], "gaussian"), (33, [22, 10, 1], [5, 5, 6], "rastrigin"), (40, [8, 25, 7], [5, 3, 6], "rastrigin"),
                            (33, [33], [16], "rastrigin"), (32, [22, 9, 1], [5, 5, 6], "rastrigin"), (33, [32, 1], [5, 6], "gaussian"), (33, [1, 32], [5, 6], "gaussian")]:
    R = sum(reps)
    rng = np.random.default_rng(7)
    kw = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D) if like == "rastrigin" else {}
    so = oracle.make_settings(D, 0, nlive=10, num_repeats=R, seed=3)
    sg = gpu.make_settings(D, 0, nlive=10, num_repeats=R, seed=3)
    cubes = np.clip(0.5 + 0.02 * rng.standard_normal((8, D)), 1e-6, 1 - 1e-6)
    chol = np.tril(rng.standard_normal((D, D)) * 0.002) + 0.01 * np.eye(D)
    if len(dims) > 1:
        gpu.set_grades(dims, reps); oracle.set_grades(dims, reps)
    rec, _ = oracle.calculate_points(so, cubes, like=like, **kw)
    logL = rec[:, -1] - 5.0
    uid = np.arange(8, dtype=np.uint64) + 100
    babies, nlike = gpu.slice_chains(sg, rec, chol, logL, uid, like=like, **kw)
    worst = 0.0; nl_ok = True
    for c in range(8):
        want, nl = oracle.slice_chain(so, rec[c], chol, float(logL[c]), int(uid[c]), like=like, **kw)
        nl_ok = nl_ok and nl == nlike[c]
        worst = max(worst, float(np.abs(babies[c][:, :D] - want[:, :D]).max()))
    gpu.set_grades(); oracle.set_grades()
    print(D, dims, reps, like, "nlike equal", nl_ok, "max |cube diff|", worst, flush=True)
