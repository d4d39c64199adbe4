"""Round-2 probe A: phase counters of the round-1 kernel at the BASELINE sizes, K = n/4 and n/2 (PC_DEBUG=1 prints dbg[])."""
import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as capi

def run(tag, settings, K=0, reps=2, **kw):
    capi.set_option("batch_K", K)
    try:
        for r in range(reps):
            t0 = time.perf_counter()
            info, _ = capi.run(settings, **kw)
            t = time.perf_counter() - t0
        d = info.as_dict()
        print(f"== {tag} K={info.batch_K} wall {t*1e3:.2f} ms device {info.device_ms:.3f} ms logZ {info.logZ:.4f}+-{info.logZerr:.4f} ndead {info.ndead} nlike {info.nlike} gens {info.ngenerations} upd {info.nupdates} evals/s {info.nlike/(info.device_ms*1e-3):.3e}")
        print("   phase_ms [wait,S,fin,U,prep,white,slice,total]:", {k: round(v, 3) for k, v in d["phase_ms"].items()}, flush=True)
    finally:
        capi.set_option("batch_K", 0)

g = lambda n, seed=1: capi.make_settings(20, 2, nlive=n, num_repeats=40, seed=seed)
run("G20 n=1000", g(1000), 250)
run("G20 n=1000", g(1000), 500)
run("G20 n=1000", g(1000), 333)
run("G20 n=8000", g(8000), 2000)
run("G20 n=8000", g(8000), 4000)
box = dict(prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
run("R10 n=2000 clustered", capi.make_settings(10, 0, nlive=2000, num_repeats=50, seed=1, do_clustering=True), 500, like="rastrigin", **box)
run("R10 n=2000 unclustered", capi.make_settings(10, 0, nlive=2000, num_repeats=50, seed=1, do_clustering=False), 500, like="rastrigin", **box)
rng = np.random.default_rng(0); D = 50
Q, _ = np.linalg.qr(rng.standard_normal((D, D))); sig = 0.1 * 0.01 ** (np.arange(D) / (D - 1))
invcov = (Q / sig ** 2) @ Q.T
params = np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]])
run("C50 n=4000 R=250", capi.make_settings(D, 0, nlive=4000, num_repeats=250, seed=1), 1000, reps=1, like="corr_gaussian", like_params=params)
