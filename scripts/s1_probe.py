import sys
sys.path.insert(0, ".")
from polychordlite_b200 import _capi as capi
n = int(sys.argv[1])
for i in range(2):
    info, _ = capi.run(capi.make_settings(20, 2, nlive=n, num_repeats=40, seed=i))
print(n, info.ngenerations, info.device_ms, {k: round(v, 3) for k, v in info.as_dict()["phase_ms"].items()})
