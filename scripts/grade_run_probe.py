import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as gpu
import oracle_lib as oracle
D = 33
kw = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D)
for grades in [None, ([22, 10, 1], [5, 5, 6]), ([22, 11], [8, 8])]:
    for like in ["rastrigin", "gaussian"]:
        for md in [118, 236, 600, 1500, -1]:
            st = dict(nlive=120, num_repeats=16, seed=710, max_ndead=md, precision_criterion=1e-2)
            k2 = kw if like == "rastrigin" else {}
            if grades:
                gpu.set_grades(*grades); oracle.set_grades(*grades)
            gpu.set_option("batch_K", 59)
            gi, _ = gpu.run(gpu.make_settings(D, 0, **st), like=like, **k2)
            gpu.set_option("batch_K", 0)
            oi, _ = oracle.run(oracle.make_settings(D, 0, batch_K=59, **st), like=like, **k2)
            gpu.set_grades(); oracle.set_grades()
            ok = (gi.ndead, gi.nlike, gi.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
            print(f"grades={grades} {like:9s} max_ndead={md}: gpu {(gi.ndead, gi.nlike, gi.nupdates, gi.nfailures)} oracle {(oi.ndead, oi.nlike, oi.nupdates, oi.nfailures)} {'OK' if ok else 'DIFF'}", flush=True)
