"""do_clustering on the device (csrc/pc_cluster.cuh): at every update every cluster
is searched for sub-clusters (device k-nearest neighbours + host union-find = the reference's NN_clustering) and split,
a cluster without live points is deleted, the phantoms take the label of their nearest live point, every cluster gets
its own covariance / Cholesky factor and a chain whitens with the factor of its seed's cluster; every death is
attributed to the cluster of the dying point (local evidences).  Checked against the oracle's batched schedule with clustering on
(same labels => same factors => same chains: identical ndead / nlike, logZ to rounding) and against the analytic
Rastrigin evidence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BOX2 = dict(prior_lo=[-5.12] * 2, prior_hi=[5.12] * 2)


@pytest.mark.parametrize("case", ["two_blobs", "three_blobs_5d", "one_blob", "ring_and_blob", "tiny", "grid_of_modes", "uniform"])
def test_device_clustering_matches_oracle_nn_clustering(gpu, oracle, case):
    rng = np.random.default_rng(5)
    if case == "two_blobs":
        pts = np.vstack([0.3 + 0.02 * rng.standard_normal((60, 2)), 0.7 + 0.02 * rng.standard_normal((40, 2))])
    elif case == "three_blobs_5d":
        pts = np.vstack([c + 0.01 * rng.standard_normal((35, 5)) for c in (0.2, 0.5, 0.8)])
    elif case == "one_blob":
        pts = 0.5 + 0.05 * rng.standard_normal((120, 4))
    elif case == "ring_and_blob":
        a = rng.uniform(0, 2 * np.pi, 80)
        pts = np.vstack([0.5 + 0.3 * np.column_stack([np.cos(a), np.sin(a)]) + 0.003 * rng.standard_normal((80, 2)),
                         0.5 + 0.01 * rng.standard_normal((30, 2))])
    elif case == "grid_of_modes":    # what a Rastrigin run looks like half-way: many small modes
        c = np.array([(i, j) for i in range(5) for j in range(5)]) / 5.0 + 0.1
        pts = np.vstack([ci + 0.004 * rng.standard_normal((rng.integers(2, 14), 2)) for ci in c])
    elif case == "uniform":
        pts = rng.uniform(0, 1, (400, 3))
    else:
        pts = rng.uniform(0, 1, (3, 3))
    pts = pts[rng.permutation(len(pts))]
    lab, num = gpu.cluster_points(pts)
    olab, onum = oracle.nn_clustering(pts)
    assert num == onum and np.array_equal(lab, olab)


@pytest.mark.parametrize("seed", [3, 4])
def test_clustered_rastrigin_run_matches_oracle(gpu, oracle, seed):
    n, R, K = 200, 6, 50
    gpu.set_option("batch_K", K)
    try:
        gi, _ = gpu.run(gpu.make_settings(2, 0, nlive=n, num_repeats=R, seed=seed, do_clustering=True), like="rastrigin", **BOX2)
    finally:
        gpu.set_option("batch_K", 0)
    oi, _ = oracle.run(oracle.make_settings(2, 0, nlive=n, num_repeats=R, seed=seed, do_clustering=True, batch_K=K),
                       like="rastrigin", **BOX2)
    assert gi.ncluster_max > 1 and oi.nsplits > 0                  # the modes were found
    assert gi.ncluster_updates == gi.nupdates
    assert (gi.ndead, gi.nlike, gi.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
    assert abs(gi.logZ - oi.logZ) < 1e-7
    assert gi.ncluster_max == oi.ncluster
    # the clusters themselves: the same identities alive and deleted, the same dead points in each, the same local
    # evidences log<Z_p> and second moments log<Z_p^2>; the local evidences add up to the global one
    g_act, g_rows, g_uid = gpu.last_clusters()
    o_act, o_rows = oracle.last_clusters()
    o_dead, o_par, o_uid = oracle.last_dead_clusters()
    g_dead, g_par = gpu.last_dead_clusters()
    assert g_act == o_act and np.array_equal(g_uid[:g_act], o_uid[:o_act])
    assert sorted(g_uid[g_act:]) == sorted(o_uid[o_act:]) and len(g_uid) == len(set(g_uid))
    assert np.array_equal(g_par, o_par) and np.array_equal(g_dead, o_dead)
    go, oo = np.argsort(g_uid), np.argsort(o_uid)
    assert np.allclose(g_rows[go], o_rows[oo][:, :2], rtol=0, atol=1e-7)
    lz = g_rows[:, 0]
    assert abs(lz.max() + np.log(np.exp(lz - lz.max()).sum()) - gi.logZ_raw) < 1e-9
    assert np.all(g_rows[:, 1] >= 2 * g_rows[:, 0] - 1e-9)          # <Z_p^2> >= <Z_p>^2


def test_unimodal_run_with_clustering_on_matches_oracle(gpu, oracle):
    gpu.set_option("batch_K", 40)
    try:
        gi, _ = gpu.run(gpu.make_settings(5, 2, nlive=160, num_repeats=10, seed=2, do_clustering=True))
    finally:
        gpu.set_option("batch_K", 0)
    oi, _ = oracle.run(oracle.make_settings(5, 2, nlive=160, num_repeats=10, seed=2, do_clustering=True, batch_K=40))
    assert (gi.ndead, gi.nlike) == (oi.ndead, oi.nlike) and abs(gi.logZ - oi.logZ) < 1e-7


def test_clustered_rastrigin_logZ_matches_analytic(gpu):
    """2-D Rastrigin as shipped (ini/rastrigin.ini), log Z = -2 ln 10.24, and the 10-D BASELINE shape at a reduced
    nlive: the evidence is unaffected by the clusters (they only steer the proposals)."""
    lz = [gpu.run(gpu.make_settings(2, 0, nlive=400, num_repeats=6, seed=s, do_clustering=True), like="rastrigin", **BOX2)[0].logZ
          for s in range(12)]
    se = np.std(lz, ddof=1) / np.sqrt(len(lz))
    assert abs(np.mean(lz) - (-2 * np.log(10.24))) < max(4 * se, 0.1)


def test_dumper_and_clustering_together(gpu):
    info, dumps = gpu.run(gpu.make_settings(2, 0, nlive=200, num_repeats=6, seed=1, do_clustering=True), like="rastrigin",
                          want_dump=True, **BOX2)
    assert len(dumps) == info.nupdates + 1 and dumps[-1]["live"].shape[0] == 0
    assert dumps[0]["live"].shape[0] == 200 and dumps[-1]["dead"].shape[0] == info.ndead
