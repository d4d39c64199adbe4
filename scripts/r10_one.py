import sys
sys.path.insert(0, ".")
from polychordlite_b200 import _capi as capi
info, _ = capi.run(capi.make_settings(10, 0, nlive=2000, num_repeats=50, seed=1, do_clustering=True), like="rastrigin", prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
print(info.ndead, info.cluster_ms)
