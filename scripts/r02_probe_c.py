"""Round-2 probe C: phase counters (PC_DEBUG=1 prints dbg[]) of single runs at growing nlive."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from polychordlite_b200 import _capi as capi
for n in [int(a) for a in sys.argv[1:]] or [1000, 8000]:
    for rep in range(2):
        t0 = time.perf_counter(); info, _ = capi.run(capi.make_settings(20, 2, nlive=n, num_repeats=40, seed=1)); t = time.perf_counter() - t0
    d = info.as_dict()
    print(f"== G20 n={n} K={info.batch_K} device {info.device_ms:.3f} ms ndead {info.ndead} nlike {info.nlike} gens {info.ngenerations} upd {info.nupdates} evals/s {info.nlike/(info.device_ms*1e-3):.3e}")
    print("   phase_ms:", {k: round(v, 3) for k, v in d["phase_ms"].items()}, flush=True)
