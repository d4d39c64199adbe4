import sys, time
sys.path.insert(0, '.')
import numpy as np
from polychordlite_b200 import _capi as capi
for want in (False, True, False, True):
    s = capi.make_settings(20, 2, nlive=1000, num_repeats=40, seed=3)
    t0 = time.perf_counter()
    info, dumps = capi.run(s, want_dump=want)
    t = time.perf_counter() - t0
    print(want, 'wall %.1f ms' % (t * 1e3), 'device_ms %.2f' % info.device_ms, 'wall_ms %.2f' % info.wall_ms,
          'launches', info.kernel_launches, 'ngen', info.ngenerations, 'nupd', info.nupdates, 'ndead', info.ndead,
          'logZ %.4f' % info.logZ, 'dumps', len(dumps), flush=True)
    print('   ', {k: round(v, 2) for k, v in info.as_dict()['phase_ms'].items()}, flush=True)
