"""One device-resident run of a BASELINE configuration (target of ncu; PC_DEBUG=1 prints the phase counters):
    python scripts/r02_cfg_once.py C3|C4|C2 [batch_K] [max_ndead]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as capi
which = sys.argv[1] if len(sys.argv) > 1 else "C4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mx = int(sys.argv[3]) if len(sys.argv) > 3 else -1
capi.set_option("batch_K", K)
kw = {}
if which == "C4":
    D = 50
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    sig = float(np.float32(0.1)) * 0.01 ** (np.arange(D) / (D - 1))
    invcov = (Q / sig ** 2) @ Q.T
    kw = dict(like="corr_gaussian", like_params=np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]]))
    s = capi.make_settings(D, 0, nlive=4000, num_repeats=250, seed=1, max_ndead=mx)
elif which == "C3":
    kw = dict(like="rastrigin", prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
    s = capi.make_settings(10, 0, nlive=2000, num_repeats=50, seed=1, do_clustering=True, max_ndead=mx)
else:
    s = capi.make_settings(20, 2, nlive=1000, num_repeats=40, seed=1, max_ndead=mx)
info, _ = capi.run(s, **kw)
d = info.as_dict()
print(which, "K", info.batch_K, "device_ms", round(info.device_ms, 3), "ndead", info.ndead, "nlike", info.nlike, "gens", info.ngenerations,
      "updates", info.nupdates, "logZ", info.logZ, "kernel", d.get("kernel"))
print("phase_ms", {k: round(v, 3) for k, v in d["phase_ms"].items()})
