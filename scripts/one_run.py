"""One device-resident run of the bench workload (for ncu --set full)."""
import sys
sys.path.insert(0, ".")
from polychordlite_b200 import _capi as capi
info, _ = capi.run(capi.make_settings(20, 2, nlive=1000, num_repeats=40, seed=0))
print(info.as_dict())
