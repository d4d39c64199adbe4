"""One device-resident run of the bench workload (for ncu --set full; PC_DEBUG=1 prints the phase counters)."""
import sys
sys.path.insert(0, ".")
from polychordlite_b200 import _capi as capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for i in range(n):
    info, _ = capi.run(capi.make_settings(20, 2, nlive=1000, num_repeats=40, seed=i))
print(info.as_dict())
