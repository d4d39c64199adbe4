import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from polychordlite_b200 import _capi as gpu
import oracle_lib as oracle
D = 33
kw = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D)
grades = ([22, 11], [8, 8])
st = dict(nlive=120, num_repeats=16, seed=710, max_ndead=-1, precision_criterion=1e-2)
gpu.set_grades(*grades); oracle.set_grades(*grades)
gpu.set_option("batch_K", 59)
gi, gd = gpu.run(gpu.make_settings(D, 0, **st), like="rastrigin", want_dump=True, **kw)
oi, od = oracle.run(oracle.make_settings(D, 0, batch_K=59, **st), like="rastrigin", want_dump=True, **kw)
g, o = gd[-1]["dead"], od[-1]["dead"]
m = min(len(g), len(o))
diff = np.abs(g[:m] - o[:m]).max(axis=1)
bad = np.flatnonzero(diff > 1e-9)
print("first differing dead row", bad[:5], "of", len(g), len(o))
if len(bad):
    i = bad[0]
    print("gen of first diff ~", i // 59, "row diff", diff[i], "logL gpu/oracle", g[i, -1], o[i, -1], "birth", g[i, -2], o[i, -2])
    j = max(0, i - 3)
    print("max diff of rows before", diff[:i].max())
    print("logL around (gpu):", g[j:i + 3, -1]); print("logL around (ora):", o[j:i + 3, -1])
    print("theta gpu", g[i, :6]); print("theta ora", o[i, :6])
    # which dims differ
    print("dims differing", np.flatnonzero(np.abs(g[i, :D] - o[i, :D]) > 1e-12))
