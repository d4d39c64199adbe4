"""Cluster outputs of a run with do_clustering: the "Local evidences" table of
<root>.stats (read_write.F90:858-872) read back by the mirror of the reference's parser (pypolychord/output.py:57-99), and
the cluster posterior files clusters/<root>_<i>.txt / _equal_weights.txt (read_write.F90:527-607)."""
import numpy as np
import pytest

from polychordlite_b200 import pypolychord
from polychordlite_b200.pypolychord import PolyChordSettings
from polychordlite_b200.pypolychord.builtin import Rastrigin
from polychordlite_b200.pypolychord.priors import UniformPrior

pytestmark = pytest.mark.gpu


def test_stats_table_and_cluster_posterior_files(gpu, tmp_path):
    s = PolyChordSettings(2, 0, nlive=300, num_repeats=6, feedback=0, do_clustering=True, write_resume=False, read_resume=False,
                          base_dir=str(tmp_path), file_root="r2", seed=3, cluster_posteriors=True, posteriors=True, equals=True,
                          write_live=False, write_prior=False)
    out = pypolychord.run_polychord(Rastrigin(), 2, 0, s, UniformPrior(-5.12, 5.12))
    nact, rows, uid = gpu.last_clusters()
    info = gpu.last_run_info()
    assert len(uid) > 3 and out.ncluster == len(uid) and info.ncluster_max > 1
    # the table: calculate_logZ_estimate per cluster (run_time_info.f90:652-678), every cluster listed
    assert np.allclose(out.logZs, 2 * rows[:, 0] - 0.5 * rows[:, 1], rtol=0, atol=1e-9)
    assert np.allclose(out.logZerrs, np.sqrt(np.abs(rows[:, 1] - 2 * rows[:, 0])), rtol=0, atol=1e-9)
    text = (tmp_path / "r2.stats").read_text()
    assert f" ncluster:   {0:8d} /{len(uid):8d}" in text          # every live point died in the final kill-off
    assert "log(Z_1)     = " in text and (len(uid) < 10 or "log(Z_10)    = " in text)
    # the local evidences add up to the global one
    lz = rows[:, 0]
    assert abs(lz.max() + np.log(np.exp(lz - lz.max()).sum()) - info.logZ_raw) < 1e-9
    # cluster files: one pair per cluster, ordered by local evidence; the points lie inside the prior box, the largest
    # cluster's files carry its share of the evidence as the largest weight
    cdir = tmp_path / "clusters"
    assert len(sorted(cdir.glob("r2_*.txt"))) == 2 * len(uid)
    first = np.loadtxt(cdir / "r2_1.txt", ndmin=2)
    share = np.exp(np.sort(lz)[::-1] - info.logZ_raw)
    assert first.shape[1] == 2 + 2 and np.all(np.abs(first[:, 2:]) <= 5.12)
    assert abs(first[:, 0].max() - share[0]) < 1e-9
    eq = np.loadtxt(cdir / "r2_1_equal_weights.txt", ndmin=2)
    assert eq.shape[0] > 0 and np.allclose(eq[:, 0], share[0])
    dead = np.loadtxt(tmp_path / "r2_dead.txt", ndmin=2)              # logL, theta
    assert set(map(tuple, np.round(first[:, 2:], 9))) <= set(map(tuple, np.round(dead[:, 1:3], 9)))
