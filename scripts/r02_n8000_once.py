"""The 8-GPU BASELINE problem (20-D Gaussian, nlive 8000) on ONE GPU, for the strong-scaling comparison."""
import sys
sys.path.insert(0, ".")
from polychordlite_b200 import _capi as capi
for i in range(2):
    info, _ = capi.run(capi.make_settings(20, 2, nlive=8000, num_repeats=40, seed=i))
d = info.as_dict()
print("G20 n=8000 one GPU: K", info.batch_K, "device_ms", round(info.device_ms, 3), "ndead", info.ndead, "nlike", info.nlike, "logZ", info.logZ)
print("phase_ms", {k: round(v, 3) for k, v in d["phase_ms"].items()})
