"""Phase U's record stream (csrc/pc_run_kernel.cuh phase_UB): with an even record length the kept phantoms move by bulk
copies (cp.async.bulk + mbarriers) through a per-warp staging ring; the register path stays for odd record lengths and
behind the option "no_bulk".  Both must compact the same records into the same places and add the same moments in the same
order: the runs are bit-identical.  Reference: clean_phantoms + calculate_covmats, run_time_info.f90:601-641, 820-877."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(gpu, no_bulk, like="gaussian", extra=None, K=0, **kw):
    gpu.set_option("no_bulk", no_bulk)
    gpu.set_option("batch_K", K)
    try:
        return gpu.run(gpu.make_settings(**kw), like=like, want_dump=True, **(extra or {}))
    finally:
        gpu.set_option("no_bulk", 0)
        gpu.set_option("batch_K", 0)


@pytest.mark.parametrize("kw,like,extra", [
    (dict(nDims=20, nDerived=2, nlive=1000, num_repeats=40, seed=3), "gaussian", None),               # BASELINE config 2: T = 44
    (dict(nDims=10, nDerived=0, nlive=400, num_repeats=20, seed=4, do_clustering=True), "rastrigin",
     dict(prior_lo=[-5.12] * 10, prior_hi=[5.12] * 10)),                                                # labels move with the records
    (dict(nDims=33, nDerived=0, nlive=200, num_repeats=40, seed=5, max_ndead=1500), "gaussian", None),  # two moment passes (D > 31)
])
def test_bulk_stream_equals_register_stream(gpu, kw, like, extra):
    a, da = run(gpu, 0, like, extra, **kw)
    b, db = run(gpu, 1, like, extra, **kw)
    assert (a.ndead, a.nlike, a.nupdates, a.ngenerations) == (b.ndead, b.nlike, b.nupdates, b.ngenerations)
    assert a.logZ == b.logZ and a.logZerr == b.logZerr
    assert len(da) == len(db)
    assert np.array_equal(da[-1]["dead"], db[-1]["dead"]) and np.array_equal(da[-1]["logweights"], db[-1]["logweights"])


def test_odd_record_length_takes_the_register_stream_and_matches_oracle(gpu, oracle):
    """nDerived = 1: T = 2 D + 3 is odd, records are not 16-byte aligned, no bulk copies."""
    kw = dict(nDims=6, nDerived=1, nlive=120, num_repeats=12, seed=8)
    gpu.set_option("batch_K", 30)
    try:
        gi, _ = gpu.run(gpu.make_settings(**kw))
    finally:
        gpu.set_option("batch_K", 0)
    oi, _ = oracle.run(oracle.make_settings(batch_K=30, **kw))
    assert (gi.ndead, gi.nlike, gi.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
    assert abs(gi.logZ - oi.logZ) < 1e-7


def test_automatic_batch_size_is_whole_waves_of_chains(gpu):
    """a run alone on the device takes m waves of (SMs - 1) * 4 chains per generation when that lies in
    [0.5, 0.6] nlive, else nlive / 2; a sharded run counts the waves of all its devices; an explicit batch_K wins."""
    import torch
    wave = (torch.cuda.get_device_properties(0).multi_processor_count - 1) * 4
    for n in (200, 500, 1000, 2000, 8000):
        K = gpu.auto_batch_size(n, 1)
        m = -(-(n // 2) // wave)
        assert K == (m * wave if m * wave <= 0.6 * n else round(n / 2)), (n, K)
    assert gpu.auto_batch_size(8000, 8) == gpu.auto_batch_size(1000, 1) * 8 or gpu.auto_batch_size(1000, 1) == 500
    info, _ = gpu.run(gpu.make_settings(20, 2, nlive=1000, num_repeats=40, seed=2))
    assert info.batch_K == gpu.auto_batch_size(1000, 1)
    gpu.set_option("batch_K", 123)
    try:
        assert gpu.auto_batch_size(1000, 1) == 123
    finally:
        gpu.set_option("batch_K", 0)


def test_narrow_phantom_traffic_changes_nothing_in_an_ensemble(gpu):
    """ChainParams::ph_narrow (ensembles): the chains do not store a phantom's theta and phase U does not carry it along.
    Nothing reads it, so every run of the ensemble ends with the same numbers as with full records."""
    s = gpu.make_settings(20, 2, nlive=300, num_repeats=40)
    seeds = list(range(6))
    a = gpu.run_ensemble(s, seeds)
    gpu.set_option("no_narrow", 1)
    try:
        b = gpu.run_ensemble(s, seeds)
    finally:
        gpu.set_option("no_narrow", 0)
    for x, y in zip(a, b):
        assert (x.ndead, x.nlike, x.nupdates) == (y.ndead, y.nlike, y.nupdates)
        assert x.logZ == y.logZ and x.logZerr == y.logZerr
