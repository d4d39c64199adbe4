"""The BASELINE.json configurations at FULL size (C3 Rastrigin 10-D nlive=2000 clustered,
C4 50-D correlated Gaussian nlive=4000 R=250, C5's nlive=8000 on one GPU): run parity with the oracle's batched
schedule over a bounded number of deaths (max_ndead keeps the CPU side to seconds; the schedule up to there is the
complete run's), and ensemble statistics of complete runs against the analytic evidences."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def corr50_params():
    """random_gaussian.f90: mu = 0.5, sigma_j = 0.1 * 0.01^((j-1)/(D-1)) (sigma a single-precision literal), Haar basis."""
    D = 50
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    sig = float(np.float32(0.1)) * 0.01 ** (np.arange(D) / (D - 1))
    invcov = (Q / sig ** 2) @ Q.T
    return np.concatenate([np.full(D, 0.5), invcov.flatten(order="F"), [2 * np.log(sig).sum()]])


def parity(gpu, oracle, K, like="gaussian", extra=None, **kw):
    extra = extra or {}
    gpu.set_option("batch_K", K)
    try:
        gi, _ = gpu.run(gpu.make_settings(**kw), like=like, **extra)
    finally:
        gpu.set_option("batch_K", 0)
    oi, _ = oracle.run(oracle.make_settings(batch_K=K, **kw), like=like, **extra)
    assert (gi.ndead, gi.nlike, gi.nchains, gi.ngenerations, gi.nupdates, gi.nfailures) == \
           (oi.ndead, oi.nlike, oi.nchains, oi.ngenerations, oi.nupdates, oi.nfailures)
    assert abs(gi.logZ - oi.logZ) < 1e-6
    return gi, oi


def test_rastrigin10_nlive2000_clustered_matches_oracle_over_12n_deaths(gpu, oracle):
    """BASELINE config 3 at full size, the first 12 nlive deaths (clustering passes included)."""
    box = dict(prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
    gi, oi = parity(gpu, oracle, 1000, like="rastrigin", extra=box, nDims=10, nDerived=0, nlive=2000, num_repeats=50, seed=1,
                    do_clustering=True, max_ndead=24000)
    assert gi.ncluster_max == oi.ncluster and gi.ncluster_max > 1


def test_rastrigin10_nlive2000_clustered_ensemble(gpu):
    """Complete runs of config 3, 16 seeds, against the analytic -10 ln 10.24 and against what the ORACLE does on the same
    problem (tests/golden/c3_scatter_oracle.json, scripts/r02_c3_scatter.py): in the reference schedule with per-cluster
    evidences the run-to-run scatter of log Z is 0.55 -- four times the error bar the runs report -- and 0.60 in the
    batched schedule with the evidence kept global.  The excess belongs to slice sampling this likelihood (11^10 modes, a
    handful of clusters found), so the engine is held to the oracle's scatter, not to the reported error bar."""
    import json
    from pathlib import Path
    gold = json.loads((Path(__file__).parent / "golden" / "c3_scatter_oracle.json").read_text())
    ref = gold["reference_schedule_per_cluster_evidence"]
    box = dict(prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)
    z, e = [], []
    for seed in range(16):
        info, _ = gpu.run(gpu.make_settings(10, 0, nlive=2000, num_repeats=50, seed=seed, do_clustering=True), like="rastrigin", **box)
        z.append(info.logZ); e.append(info.logZerr)
    z, e = np.array(z), np.array(e)
    sd = z.std(ddof=1)
    assert abs(z.mean() - gold["logZ_true"]) < 0.1 + 3 * sd / np.sqrt(len(z)), (z.mean(), sd, e.mean())
    # F test against the oracle's 8 reference-schedule runs: variances within a factor 3.5 (99 % for 15 / 7 degrees of freedom)
    assert sd ** 2 < 3.5 * ref["std"] ** 2, (sd, ref["std"])
    assert 0.08 < e.mean() < 0.25


def test_corr_gaussian50_nlive4000_matches_oracle_bounded(gpu, oracle):
    """BASELINE config 4 at full size (R = 250 slice steps per chain, directions in global scratch), the first 3000 deaths."""
    lp = corr50_params()
    parity(gpu, oracle, 1000, like="corr_gaussian", extra=dict(like_params=lp), nDims=50, nDerived=0, nlive=4000,
           num_repeats=250, seed=2, max_ndead=3000)


def test_corr_gaussian50_nlive4000_evidence(gpu):
    """Complete runs of config 4: log Z = 0 (the Gaussian sits 5 sigma inside the cube), sqrt(H/n) = 0.2 per run."""
    lp = corr50_params()
    z, e = [], []
    for seed in range(4):
        info, _ = gpu.run(gpu.make_settings(50, 0, nlive=4000, num_repeats=250, seed=seed), like="corr_gaussian", like_params=lp)
        z.append(info.logZ); e.append(info.logZerr)
    z, e = np.array(z), np.array(e)
    assert abs(z.mean()) < 4 * e.mean() / np.sqrt(len(z)) + 0.05, (z, e)
    assert np.all(e > 0.1) and np.all(e < 0.4)


def test_gaussian20_nlive8000_on_one_gpu_matches_oracle_bounded(gpu, oracle):
    """BASELINE config 5's live set on ONE device (the sharded form: tests/test_gpu_sharded.py, bench.py --gpus N)."""
    parity(gpu, oracle, 4000, nDims=20, nDerived=2, nlive=8000, num_repeats=40, seed=3, max_ndead=16000)
