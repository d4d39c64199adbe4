"""Round-2 probe B: ensemble throughput of the dense chain phase against the warp-per-chain phase, and the phase
counters of single runs (PC_DEBUG=1 prints dbg[])."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as capi

def ens(tag, nruns, dense, K=0, n=1000):
    capi.set_option("dense", dense); capi.set_option("batch_K", K)
    try:
        s = capi.make_settings(20, 2, nlive=n, num_repeats=40)
        for rep in range(2):
            t0 = time.perf_counter()
            infos = capi.run_ensemble(s, list(range(nruns)))
            t = time.perf_counter() - t0
        nl = sum(i.nlike for i in infos); z = np.array([i.logZ for i in infos])
        dms = infos[0].device_ms
        print(f"== ensemble {tag} runs={nruns} dense={dense} K={infos[0].batch_K} ctas/run={infos[0].ctas_per_run} W={infos[0].warps_per_cta} "
              f"device {dms:.2f} ms wall {t*1e3:.1f} ms evals/s {nl/(dms*1e-3):.3e} logZ {z.mean():.4f}+-{z.std(ddof=1)/np.sqrt(nruns):.4f} "
              f"algGB/s {sum(i.algorithmic_bytes for i in infos)/(dms*1e-3)/1e9:.1f}", flush=True)
    finally:
        capi.set_option("dense", 0); capi.set_option("batch_K", 0)

def run(tag, settings, K=0, reps=2, **kw):
    capi.set_option("batch_K", K)
    try:
        for r in range(reps):
            t0 = time.perf_counter(); info, _ = capi.run(settings, **kw); t = time.perf_counter() - t0
        d = info.as_dict()
        print(f"== {tag} K={info.batch_K} wall {t*1e3:.2f} ms device {info.device_ms:.3f} ms logZ {info.logZ:.4f}+-{info.logZerr:.4f} ndead {info.ndead} nlike {info.nlike} gens {info.ngenerations} upd {info.nupdates} evals/s {info.nlike/(info.device_ms*1e-3):.3e}")
        print("   phase_ms:", {k: round(v, 3) for k, v in d["phase_ms"].items()}, flush=True)
    finally:
        capi.set_option("batch_K", 0)

which = sys.argv[1:] or ["ens", "single"]
if "ens" in which:
    ens("spec", 32, -1)
    ens("dense", 32, 0)
    ens("dense", 64, 0)
    ens("dense", 72, 0)
    ens("dense K=500", 72, 0, K=500)
if "single" in which:
    g = lambda n, seed=1: capi.make_settings(20, 2, nlive=n, num_repeats=40, seed=seed)
    run("G20 n=1000", g(1000), 250)
    run("G20 n=1000 auto", g(1000), 0)
    run("G20 n=8000 auto", g(8000), 0)
    box = dict(prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
    run("R10 n=2000 clustered", capi.make_settings(10, 0, nlive=2000, num_repeats=50, seed=1, do_clustering=True), 0, like="rastrigin", **box)
    rng = np.random.default_rng(0); D = 50
    Q, _ = np.linalg.qr(rng.standard_normal((D, D))); sig = 0.1 * 0.01 ** (np.arange(D) / (D - 1))
    invcov = (Q / sig ** 2) @ Q.T
    params = np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]])
    run("C50 n=4000 R=250", capi.make_settings(D, 0, nlive=4000, num_repeats=250, seed=1), 0, reps=1, like="corr_gaussian", like_params=params)
