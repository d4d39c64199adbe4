"""Round-2 study (CPU oracle only): does keeping the evidence PER CLUSTER, as the reference does (run_time_info.f90:211-296,
303-598: local volumes that shrink with the cluster's own deaths, seeds drawn by volume), estimate log Z as well as keeping
it GLOBAL and attributing every death to its cluster?  2-D Rastrigin as shipped (ini/rastrigin.ini: box +-5.12,
log Z = -2 ln 10.24), nlive 400, num_repeats 6, 24 seeds per row:

  reference schedule (one death per iteration), no clustering / clustering with per-cluster evidence
  batched schedule K = 10, 50, 100: no clustering / clusters + global evidence (do_clustering=1, the engine's) /
                                    clusters + per-cluster evidence (do_clustering=2)

Writes tests/golden/cluster_bias_oracle.json.  Run:  python scripts/r02_cluster_bias.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as oracle  # noqa: E402

TRUE = -2 * np.log(10.24)
rows = []
for mode, name in ((0, "no clustering"), (1, "clusters, global evidence, deaths attributed"), (2, "clusters, per-cluster evidence")):
    for K in (0, 10, 50, 100):
        if mode == 1 and K == 0:
            continue   # the reference schedule with clustering IS per-cluster evidence (the mode-2 row)
        z, e = [], []
        for seed in range(24):
            s = oracle.make_settings(2, 0, nlive=400, num_repeats=6, seed=100 + seed, do_clustering=mode, batch_K=K)
            r, _ = oracle.run(s, like="rastrigin", prior_lo=[-5.12] * 2, prior_hi=[5.12] * 2)
            z.append(r.logZ); e.append(r.logZerr)
        z = np.array(z)
        rows.append(dict(evidence=name, schedule="reference (1 death / iteration)" if K == 0 else f"batched K={K}",
                         mean_minus_true=round(float(z.mean() - TRUE), 4), sem=round(float(z.std(ddof=1) / np.sqrt(len(z))), 4),
                         std=round(float(z.std(ddof=1)), 4), mean_reported_logZerr=round(float(np.mean(e)), 4)))
        print(rows[-1], flush=True)
out = dict(what="log Z of the 2-D Rastrigin problem in the CPU oracle, 24 seeds per row (scripts/r02_cluster_bias.py)",
           logZ_true=round(float(TRUE), 6), rows=rows,
           reading="Without clustering every schedule is unbiased.  With clustering the estimate is high by 0.1-0.3 whichever way "
                   "the evidence is kept -- per cluster as the reference keeps it (also in the reference's own schedule), or "
                   "globally: the excess comes from chains that whiten with their cluster's factor and stay in their mode at "
                   "num_repeats = 3 nDims, not from the bookkeeping.  Keeping the evidence global is therefore no worse, is "
                   "exact in itself, and costs no per-death cross-moment updates; the engine does that and attributes the deaths.")
(ROOT / "tests" / "golden" / "cluster_bias_oracle.json").write_text(json.dumps(out, indent=1))
