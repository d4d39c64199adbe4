import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as gpu
import oracle_lib as oracle
def corr(D, rng):
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    sig = 0.1 * 0.1 ** (np.arange(D) / max(D - 1, 1))
    invcov = (Q / sig ** 2) @ Q.T
    invcov = 0.5 * (invcov + invcov.T)
    return np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]])
for D, like in [(65, "corr_gaussian"), (65, "gaussian"), (40, "gaussian"), (33, "gaussian"), (31, "gaussian"), (50, "corr_gaussian")]:
    rng = np.random.default_rng(D)
    lp = corr(D, rng) if like == "corr_gaussian" else None
    for n, R, K, md in [(16, 60, 12, 24), (16, 60, 12, 100), (16, 60, 12, 400), (100, 40, 25, 600), (100, 40, 25, -1)]:
        st = dict(nlive=n, num_repeats=R, seed=5, max_ndead=md, precision_criterion=1e-2)
        gpu.set_option("batch_K", K)
        gi, _ = gpu.run(gpu.make_settings(D, 0, **st), like=like, like_params=lp)
        gpu.set_option("batch_K", 0)
        oi, _ = oracle.run(oracle.make_settings(D, 0, batch_K=K, **st), like=like, like_params=lp)
        ok = (gi.ndead, gi.nlike, gi.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
        print(f"D={D} {like:14s} n={n} R={R} K={K} max_ndead={md}: gpu {(gi.ndead, gi.nlike, gi.nupdates)} oracle {(oi.ndead, oi.nlike, oi.nupdates)} {'OK' if ok else 'DIFF'} dlogZ={gi.logZ-oi.logZ:.2e}", flush=True)
