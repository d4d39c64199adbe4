"""The oracle's restatement of the reference's clustering (clustering.f90: NN_clustering, compute_knn,
do_clustering_k; utils.F90 relabel; run_time_info.f90 add_cluster / delete_cluster) -- row a19 of SURVEY.md section 8.

The restatement follows the Fortran loop by loop (pairwise label merging, insertion-sorted neighbour lists,
recursion on sub-matrices).  It is pinned here to an independent formulation of the same definition -- connected
components of the "n mutual-nearest-neighbour" graph by union-find, canonical labels by first appearance, a
work list instead of recursion -- and to known answers; clustered runs are pinned to the analytic evidence of the
Rastrigin function (rastrigin.f90:33)."""
import numpy as np
import pytest


def _components(m, edges):
    parent = list(range(m))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for a, b in edges:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    roots = [find(i) for i in range(m)]
    first = {}
    return np.array([first.setdefault(r, len(first)) for r in roots]), len(first)


def _nn_split(pts):
    """one NN_clustering level on the points `pts`: labels, number of clusters"""
    m = len(pts)
    if m < 2:
        return np.zeros(m, int), 1
    k = min(m, 10)
    d2 = np.zeros((m, m))
    for kk in range(pts.shape[1]):                      # same accumulation order as the oracle (dimension by dimension)
        d2 += (pts[:, kk][:, None] - pts[:, kk][None, :]) ** 2
    knn = np.argsort(d2, axis=1, kind="stable")[:, :k]  # ties keep index order (compute_knn's insertion rule)
    old = np.arange(m)
    lab, num = old, m
    for n in range(2, k + 1):
        heads = knn[:, 0]
        edges = [(i, j) for i in range(m) for j in range(m) if i < j and (heads[j] in knn[i, :n] or heads[i] in knn[j, :n])]
        lab, num = _components(m, edges)
        if num == 1 or np.array_equal(lab, old):
            break
        old = lab
    return lab, num


def brute_nn_clustering(pts):
    pts = np.asarray(pts)
    m = len(pts)
    label = np.zeros(m, int)
    work, final, nxt = [0], set(), 1
    while work:
        c = work.pop()
        idx = np.flatnonzero(label == c)
        lab, num = _nn_split(pts[idx])
        if num == 1:
            final.add(c)
            continue
        for s in range(num):
            if s == 0:
                work.append(c)
            else:
                label[idx[lab == s]] = nxt
                work.append(nxt)
                nxt += 1
    first = {}
    return np.array([first.setdefault(v, len(first)) for v in label]), len(first)


@pytest.mark.parametrize("case", ["two_blobs", "three_blobs_5d", "one_blob", "ring_and_blob", "tiny"])
def test_nn_clustering_matches_an_independent_formulation(oracle, case):
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
    else:
        pts = rng.uniform(0, 1, (3, 3))
    pts = pts[rng.permutation(len(pts))]
    lab, num = oracle.nn_clustering(pts)
    blab, bnum = brute_nn_clustering(pts)
    assert num == bnum and np.array_equal(lab, blab)
    assert lab[0] == 0 and set(lab) == set(range(num))          # relabel: labels in order of first appearance
    if case == "two_blobs":
        assert num == 2
    if case == "one_blob":
        assert num == 1


@pytest.mark.parametrize("batch_K", [0, 50])
def test_clustered_rastrigin_logZ_matches_analytic(oracle, batch_K):
    """2-D Rastrigin as shipped (ini/rastrigin.ini): log Z = -2 ln 10.24; do_clustering on.  Reference schedule with
    the reference's per-cluster evidences (add_cluster / delete_cluster), and the engine's batched schedule."""
    lz, found = [], 0
    for seed in range(16):
        r, _ = oracle.run(oracle.make_settings(2, 0, nlive=200, num_repeats=6, seed=seed, do_clustering=True,
                                               batch_K=batch_K), like="rastrigin", prior_lo=[-5.12] * 2, prior_hi=[5.12] * 2)
        lz.append(r.logZ)
        found += r.nsplits
    se = np.std(lz, ddof=1) / np.sqrt(len(lz))
    assert found > 16                                           # clusters were found in every run or so
    assert abs(np.mean(lz) - (-2 * np.log(10.24))) < max(4 * se, 0.1)


def test_clustering_off_and_on_agree_statistically(oracle):
    a = [oracle.run(oracle.make_settings(3, 0, nlive=150, num_repeats=9, seed=s, do_clustering=c))[0].logZ
         for c in (False, True) for s in range(10)]
    off, on = np.array(a[:10]), np.array(a[10:])
    assert abs(off.mean() - on.mean()) < 4 * np.sqrt(off.var(ddof=1) / 10 + on.var(ddof=1) / 10) + 0.05


@pytest.mark.parametrize("mode", [1, 2])
def test_batched_clusters_keep_their_identity_and_their_evidences_add_up(oracle, mode):
    """The batched schedule with clustering (the engine's): clusters persist (split at updates, deleted when empty), every
    death belongs to one cluster, and the local evidences add up to the global one -- with the evidence kept globally and
    the deaths attributed (mode 1, the engine's) and with per-cluster volumes as the reference keeps them (mode 2)."""
    s = oracle.make_settings(2, 0, nlive=300, num_repeats=6, seed=4, do_clustering=mode, batch_K=75)
    r, _ = oracle.run(s, like="rastrigin", prior_lo=[-5.12] * 2, prior_hi=[5.12] * 2)
    nact, rows = oracle.last_clusters()
    dead, parent, uid = oracle.last_dead_clusters()
    assert r.ncluster > 1 and len(rows) == len(uid) > nact >= 1
    lz = rows[:, 0]
    assert abs(lz.max() + np.log(np.exp(lz - lz.max()).sum()) - r.logZ_raw) < 1e-9
    assert np.all(rows[:, 1] >= 2 * rows[:, 0] - 1e-9)                       # <Z_p^2> >= <Z_p>^2
    assert len(dead) == r.ndead and dead.min() >= 0 and dead.max() < len(parent)
    assert parent[0] == -1 and np.all(parent[1:] >= 0) and np.all(parent[1:] < np.arange(1, len(parent)))   # a tree, parents first
    assert len(set(uid)) == len(uid)
    # a cluster that was split is neither alive nor deleted at the end: its evidence went to its pieces
    assert not (set(parent[1:]) & set(uid))
    # every listed cluster that has dead points of its own has a finite local evidence
    own = set(dead)
    assert all(np.isfinite(rows[i, 0]) and rows[i, 0] > -1e29 for i, u in enumerate(uid) if u in own)
