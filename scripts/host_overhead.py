"""Where the host-side time of a run goes (PC_DEBUG=1 prints the engine's own marks)."""
import sys, time, ctypes as C, subprocess
sys.path.insert(0, '.')
import numpy as np
from polychordlite_b200 import _capi as capi
L = capi.lib()
def light(ndead, nlive, npars, live, dead, lw, logZ, logZerr):
    pass
dcb = capi.DUMPER_CB(light)
def go(tag, n=4, dumper=None):
    for i in range(n):
        s = capi.make_settings(20, 2, nlive=1000, num_repeats=40, seed=i)
        info = capi.RunInfo()
        t0 = time.perf_counter()
        rc = L.pc_run(C.byref(s), 0, None, 0, None, 0, dumper if dumper else C.cast(None, capi.DUMPER_CB), C.byref(info))
        t = (time.perf_counter() - t0) * 1e3
        print(f'{tag} wall {t:.2f} ms  engine wall {info.wall_ms:.2f}  device {info.device_ms:.2f}  launches {info.kernel_launches}', flush=True)
go('warm', 2)
go('pc_run          ')
go('pc_run + dumper ', dumper=dcb)
sys.path.insert(0, '.')
import bench
smp = bench.ClockSampler(0)
smp.start()
time.sleep(0.2)
go('nvml: pc_run         ')
go('nvml: pc_run + dumper', dumper=dcb)
print(smp.stop())
