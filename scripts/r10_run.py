"""One clustered Rastrigin 10-D run (BASELINE config 3); PC_DEBUG=1 prints where the clustering passes spend their time."""
import sys, time
sys.path.insert(0, ".")
from polychordlite_b200 import _capi as capi
for i in range(2):
    t0 = time.perf_counter()
    info, _ = capi.run(capi.make_settings(10, 0, nlive=2000, num_repeats=50, seed=i, do_clustering=True), like="rastrigin",
                       prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
    print("wall %.1f ms" % ((time.perf_counter() - t0) * 1e3), "device", round(info.device_ms, 2), "cluster_ms", round(info.cluster_ms, 2),
          "logZ", round(info.logZ, 3), "ndead", info.ndead, "nupd", info.nupdates, "ncl_max", info.ncluster_max, "launches", info.kernel_launches)
