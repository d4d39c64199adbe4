"""The BASELINE.json configurations at full size, one run each (sanity + timing; not the bench)."""
import sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as capi


def show(tag, info, t, truth=None):
    print(f"{tag:40s} wall {t*1e3:8.1f} ms  device {info.device_ms:8.2f} ms (+{info.cluster_ms:.1f} clustering)  logZ {info.logZ:9.3f} +- {info.logZerr:.3f}"
          + (f" (true {truth:.3f})" if truth is not None else "") + f"  ndead {info.ndead}  nlike {info.nlike}  evals/s {info.nlike/t:.3e}  gens {info.ngenerations} upd {info.nupdates} ncl {info.ncluster_max}", flush=True)


def timed(fn):
    fn()
    t0 = time.perf_counter(); r = fn(); return r, time.perf_counter() - t0


(i, _), t = timed(lambda: capi.run(capi.make_settings(20, 2, nlive=500, num_repeats=40, seed=1)))
show("C1 gaussian20 nlive=500", i, t, -1.15e-5)
(i, _), t = timed(lambda: capi.run(capi.make_settings(20, 2, nlive=1000, num_repeats=40, seed=1)))
show("C2 gaussian20 nlive=1000", i, t, -1.15e-5)
box = dict(prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
(i, _), t = timed(lambda: capi.run(capi.make_settings(10, 0, nlive=2000, num_repeats=50, seed=1, do_clustering=True), like="rastrigin", **box))
show("C3 rastrigin10 nlive=2000 clustering", i, t, -10 * np.log(10.24))
# C4: 50-D correlated Gaussian, sigma_j = 0.1 * 0.01^((j-1)/49), random orthogonal basis (random_utils.F90:581-614)
rng = np.random.default_rng(0)
D = 50
Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
sig = 0.1 * 0.01 ** (np.arange(D) / (D - 1))
invcov = (Q / sig ** 2) @ Q.T
params = np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]])
(i, _), t = timed(lambda: capi.run(capi.make_settings(D, 0, nlive=4000, num_repeats=250, seed=1), like="corr_gaussian", like_params=params))
show("C4 corr gaussian50 nlive=4000 R=250", i, t, 0.0)
(i, _), t = timed(lambda: capi.run(capi.make_settings(20, 2, nlive=8000, num_repeats=40, seed=1)))
show("C5' gaussian20 nlive=8000 (one GPU)", i, t, -1.15e-5)
