"""Round-2 probe: chains per generation against the chain-warp capacity of the launch (paired helper warps or all warps on
chains), for the runs whose generations take several waves: C4 (K = 2000) and the nlive-8000 Gaussian on one GPU."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as capi

def corr50():
    D = 50
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    sig = float(np.float32(0.1)) * 0.01 ** (np.arange(D) / (D - 1))
    invcov = (Q / sig ** 2) @ Q.T
    return np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]])

def go(tag, s, K, nopair, **kw):
    capi.set_option("batch_K", K); capi.set_option("no_pairing", nopair)
    try:
        info, _ = capi.run(s, **kw)
    finally:
        capi.set_option("batch_K", 0); capi.set_option("no_pairing", 0)
    print(f"{tag} K={info.batch_K} nopair={nopair} W={info.warps_per_cta} device {info.device_ms:.2f} ms ndead {info.ndead} nlike {info.nlike} "
          f"evals/s {info.nlike / info.device_ms * 1e3:.3e} us/death {info.device_ms * 1e3 / info.ndead:.3f} logZ {info.logZ:.4f} +- {info.logZerr:.4f}", flush=True)

which = sys.argv[1:] or ["c4", "g8000"]
if "c4" in which:
    lp = corr50()
    s = lambda: capi.make_settings(50, 0, nlive=4000, num_repeats=250, seed=1, max_ndead=120000)
    for K, nopair in ((0, 0), (2000, 1), (1176, 1), (1764, 0), (1176, 0), (588, 0)):
        go("C4", s(), K, nopair, like="corr_gaussian", like_params=lp)
if "g8000" in which:
    s = lambda: capi.make_settings(20, 2, nlive=8000, num_repeats=40, seed=1)
    for K, nopair in ((0, 0), (4000, 1), (3528, 1), (3528, 0), (2352, 1)):
        go("G20 n=8000", s(), K, nopair)
if "g1000" in which:
    s = lambda: capi.make_settings(20, 2, nlive=1000, num_repeats=40, seed=1)
    for K, nopair in ((0, 0), (500, 1), (588, 0)):
        go("G20 n=1000", s(), K, nopair)
