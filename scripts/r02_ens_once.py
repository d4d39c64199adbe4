"""One ensemble launch of the bench workload (for ncu): argv = [nruns, dense, launches]."""
import sys
sys.path.insert(0, ".")
from polychordlite_b200 import _capi as capi
nruns = int(sys.argv[1]) if len(sys.argv) > 1 else 72
capi.set_option("dense", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
    infos = capi.run_ensemble(capi.make_settings(20, 2, nlive=1000, num_repeats=40), list(range(nruns)))
print(infos[0].device_ms, sum(i.nlike for i in infos) / (infos[0].device_ms * 1e-3))
