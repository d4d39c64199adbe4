"""GPU parity, whole runs: the persistent kernel against the oracle's batched-generation mode on
identical seeds, the analytic evidences, and the reference's behavioural contracts."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
ANALYTIC = json.loads((GOLDEN / "analytic.json").read_text())


def both(gpu, oracle, like="gaussian", extra=None, **kw):
    extra = extra or {}
    K = kw.pop("batch_K")
    so = oracle.make_settings(batch_K=K, **kw)
    sg = gpu.make_settings(**kw)
    gpu.set_option("batch_K", K)
    try:
        gi, gd = gpu.run(sg, like=like, want_dump=True, **extra)
    finally:
        gpu.set_option("batch_K", 0)
    oi, od = oracle.run(so, like=like, want_dump=True, **extra)
    return gi, gd, oi, od


@pytest.mark.parametrize("kw", [
    dict(nDims=4, nDerived=1, nlive=64, num_repeats=8, seed=0, batch_K=16),
    dict(nDims=20, nDerived=2, nlive=200, num_repeats=40, seed=1, batch_K=50),
    dict(nDims=20, nDerived=2, nlive=1000, num_repeats=40, seed=2, batch_K=250),   # BASELINE config 2
    dict(nDims=20, nDerived=2, nlive=1000, num_repeats=40, seed=5, batch_K=500),   # ... at the engine's default batch (n/2: two babies per thread in phase S)
    dict(nDims=5, nDerived=0, nlive=700, num_repeats=10, seed=6, batch_K=300),     # a batch between the two register sorts
    dict(nDims=6, nDerived=0, nlive=100, num_repeats=12, seed=3, batch_K=99),      # K = nlive-1
    dict(nDims=6, nDerived=0, nlive=100, num_repeats=12, seed=3, batch_K=1),       # one death per generation
    dict(nDims=3, nDerived=0, nlive=50, num_repeats=1, seed=4, batch_K=10),        # no phantoms at all
])
def test_run_matches_oracle_batched_mode(gpu, oracle, kw):
    """Same seeds, same schedule -> same decisions: identical death/eval/update counts, and logZ,
    dead points and weights equal within 1e-7 (FP re-association; see DESIGN.md 'Parity')."""
    gi, gd, oi, od = both(gpu, oracle, **dict(kw))
    assert (gi.ndead, gi.nlike, gi.nchains, gi.ngenerations, gi.nupdates, gi.nfailures) == \
           (oi.ndead, oi.nlike, oi.nchains, oi.ngenerations, oi.nupdates, oi.nfailures)
    assert gi.nphantoms_final == oi.nphantoms_final
    assert abs(gi.logZ - oi.logZ) < 1e-7 and abs(gi.logZerr - oi.logZerr) < 1e-7
    assert len(gd) == len(od)
    for a, b in zip(gd, od):
        assert a["dead"].shape == b["dead"].shape and a["live"].shape == b["live"].shape
        assert np.allclose(a["dead"], b["dead"], rtol=0, atol=1e-6)
        assert np.allclose(a["logweights"], b["logweights"], rtol=0, atol=1e-6)
        assert abs(a["logZ"] - b["logZ"]) < 1e-7
        # live rows: the oracle holds them in slot order too
        assert np.allclose(a["live"], b["live"], rtol=0, atol=1e-6)


def test_run_matches_golden_fixture(gpu):
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())
    for case in g["runs"]:
        if case["batch_K"] == 0:
            continue
        sg = gpu.make_settings(case["D"], case["P"], nlive=case["nlive"], num_repeats=case["R"], seed=case["seed"])
        gpu.set_option("batch_K", case["batch_K"])
        try:
            gi, _ = gpu.run(sg)
        finally:
            gpu.set_option("batch_K", 0)
        assert gi.ndead == case["ndead"] and gi.nlike == case["nlike"] and gi.nupdates == case["nupdates"]
        assert abs(gi.logZ - case["logZ"]) < 1e-7


def test_rastrigin_and_box_prior_run_matches_oracle(gpu, oracle):
    D = 4
    extra = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D)
    gi, gd, oi, od = both(gpu, oracle, like="rastrigin", extra=extra, nDims=D, nDerived=0, nlive=200,
                          num_repeats=12, seed=6, batch_K=50)
    assert (gi.ndead, gi.nlike, gi.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
    assert abs(gi.logZ - oi.logZ) < 1e-7


def test_corr_gaussian_run_matches_oracle(gpu, oracle):
    D = 12
    inv, logdet = oracle.random_inverse_covmat(4, D, float(np.float32(0.1)))
    extra = dict(like_params=np.concatenate([np.full(D, 0.5), inv.ravel(order="F"), [logdet]]))
    gi, gd, oi, od = both(gpu, oracle, like="corr_gaussian", extra=extra, nDims=D, nDerived=0, nlive=150,
                          num_repeats=24, seed=8, batch_K=40)
    assert (gi.ndead, gi.nlike, gi.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
    assert abs(gi.logZ - oi.logZ) < 1e-6


def test_ensemble_logZ_gaussian20_nlive1000_matches_analytic(gpu):
    """BASELINE config 2 at full size: 20-D Gaussian, nlive=1000, num_repeats=40.  Single-run scatter
    (sqrt(H/nlive)=0.133) exceeds the north star's +-0.1, so parity is an ensemble statement:
    the mean over 32 seeds (s.e. ~0.024) must sit within 0.1 of the analytic logZ."""
    s = gpu.make_settings(20, 2, nlive=1000, num_repeats=40)
    infos = gpu.run_ensemble(s, list(range(32)))
    z = np.array([i.logZ for i in infos])
    se = z.std(ddof=1) / np.sqrt(z.size)
    assert abs(z.mean() - ANALYTIC["gaussian20_unit_cube"]["logZ"]) < 0.1
    assert abs(z.mean() - ANALYTIC["gaussian20_unit_cube"]["logZ"]) < 4 * se
    assert 0.08 < z.std(ddof=1) < 0.22
    assert all(i.nfailures == 0 for i in infos)
    e = np.mean([i.nlike / i.nslices for i in infos])
    assert 3.5 < e < 7.0


def test_ensemble_member_equals_single_run(gpu):
    """A run inside an ensemble launch is bit-identical to the same seed run alone."""
    s = gpu.make_settings(8, 1, nlive=100, num_repeats=16, seed=5)
    gpu.set_option("batch_K", 25)   # the automatic batch size differs between a run alone and an ensemble
    gpu.set_option("dense", -1)     # same chain phase on both sides (tests/test_gpu_dense.py covers the dense one)
    try:
        single, _ = gpu.run(s)
        infos = gpu.run_ensemble(s, [3, 5, 9])
    finally:
        gpu.set_option("batch_K", 0)
        gpu.set_option("dense", 0)
    assert infos[1].logZ == single.logZ and infos[1].nlike == single.nlike and infos[1].ndead == single.ndead
    assert infos[0].logZ != infos[1].logZ


def test_seed_determinism_and_nDerived_independence(gpu):
    """tests/test_run_pypolychord.py:77-119 of the reference, on the device path."""
    a, da = gpu.run(gpu.make_settings(4, 1, nlive=100, num_repeats=12, seed=2), want_dump=True)
    b, db = gpu.run(gpu.make_settings(4, 1, nlive=100, num_repeats=12, seed=2), want_dump=True)
    c, dc = gpu.run(gpu.make_settings(4, 0, nlive=100, num_repeats=12, seed=2), want_dump=True)
    d, _ = gpu.run(gpu.make_settings(4, 1, nlive=100, num_repeats=12, seed=3))
    assert a.logZ == b.logZ and a.nlike == b.nlike
    assert np.array_equal(da[-1]["dead"], db[-1]["dead"]) and np.array_equal(da[-1]["logweights"], db[-1]["logweights"])
    assert a.logZ == c.logZ and np.array_equal(da[-1]["dead"][:, :4], dc[-1]["dead"][:, :4])
    assert a.logZ != d.logZ


def test_posterior_moments_and_dumper_contract(gpu):
    D = 20
    s = gpu.make_settings(D, 2, nlive=1000, num_repeats=40, seed=4)
    info, dumps = gpu.run(s, want_dump=True)
    assert len(dumps) == info.nupdates + 1
    last = dumps[-1]
    assert last["live"].shape[0] == 0 and last["dead"].shape == (info.ndead, D + 4)
    w = np.exp(last["logweights"])
    assert np.isclose(w.sum(), 1.0, rtol=1e-10)
    theta = last["dead"][:, :D]
    mean = (w[:, None] * theta).sum(0)
    sd = np.sqrt((w[:, None] * (theta - mean) ** 2).sum(0))
    assert np.all(np.abs(mean - 0.5) < 0.02) and np.all(np.abs(sd - 0.1) < 0.015)
    assert np.all(np.diff(last["dead"][:, -1]) >= 0)
    assert np.all(last["dead"][:, -2] <= last["dead"][:, -1])
    assert dumps[0]["live"].shape == (1000, D + 4)


def test_max_ndead(gpu, oracle):
    s = gpu.make_settings(4, 0, nlive=50, num_repeats=8, max_ndead=95, seed=1)
    gpu.set_option("batch_K", 10)
    try:
        r, _ = gpu.run(s)
        r0, _ = gpu.run(gpu.make_settings(4, 0, nlive=50, num_repeats=8, max_ndead=0, seed=1))
    finally:
        gpu.set_option("batch_K", 0)
    assert r.ndead == 145 and r.nchains == 95
    assert r0.ndead == 50 and r0.nchains == 0 and r0.nlike == 50
    o, _ = oracle.run(oracle.make_settings(4, 0, nlive=50, num_repeats=8, max_ndead=95, seed=1, batch_K=10))
    assert abs(o.logZ - r.logZ) < 1e-8


def test_capacity_growth_paths(gpu):
    """Dead/phantom pools grow by relaunching the persistent kernel; results must not change."""
    s = gpu.make_settings(3, 0, nlive=40, num_repeats=6, seed=7)
    a, _ = gpu.run(s)
    gpu.set_option("cap_dead0", 1)
    gpu.set_option("cap_ph0", 1)
    try:
        b, _ = gpu.run(s)
    finally:
        gpu.set_option("cap_dead0", 0)
        gpu.set_option("cap_ph0", 0)
    assert b.kernel_launches > a.kernel_launches
    assert a.logZ == b.logZ and a.nlike == b.nlike and a.ndead == b.ndead


def test_reference_cpp_facade_drives_the_engine(gpu):
    """oracle/_ref/ref_driver: the REFERENCE'S OWN C++ facade (Settings + run_polychord, c_interface.cpp compiled
    from the reference tree by oracle/Makefile) linked against this library -- the drop-in boundary exercised
    from the reference's side.  The 20-D Gaussian's logZ must come back within 5 sigma of the analytic value."""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "ref_driver"
    if not exe.exists():
        pytest.skip("oracle/_ref/ref_driver not built (needs the reference tree at build time)")
    r = subprocess.run([str(exe), "20", "500", "3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    logZ, logZerr, ndead, nlive_last, ndumps = r.stdout.split()[-5:]
    assert abs(float(logZ) - ANALYTIC["gaussian20_unit_cube"]["logZ"]) < 5 * float(logZerr)
    assert int(ndead) > 5000 and int(nlive_last) == 0 and int(ndumps) > 5


def test_reference_cc_driver_linked_with_libchord_alone_runs(gpu, tmp_path):
    """oracle/_ref/polychord_CC: the reference's src/drivers/polychord_CC.cpp + likelihoods/CC/CC_likelihood.cpp, compiled
    unchanged and linked with -lchord only (oracle/Makefile) -- Settings, run_polychord and the engine all come from this
    repository's library.  Its likelihood is a host callback, so this is the lock-step path end to end."""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "polychord_CC"
    if not exe.exists():
        pytest.skip("oracle/_ref/polychord_CC not built (needs the reference tree at build time)")
    (tmp_path / "chains" / "clusters").mkdir(parents=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "log(Z)" in r.stdout
    assert (tmp_path / "chains" / "test_phys_live.txt").exists()   # settings.write_live = true (polychord_CC.cpp:24)


def test_reference_cpython_shim_linked_with_libchord_alone_runs(gpu, tmp_path):
    """oracle/_ref/refshim/_pypolychord.so: the reference's pypolychord/_pypolychord.cpp compiled unchanged against our
    library; a Python likelihood driven through it (34 positional arguments, polychord.py:600-634)."""
    import subprocess
    import sys
    from pathlib import Path
    d = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "refshim"
    if not (d / "_pypolychord.so").exists():
        pytest.skip("oracle/_ref/refshim not built (needs the reference tree at build time)")
    (tmp_path / "chains").mkdir()
    code = f"""
import sys, numpy as np
sys.path.insert(0, {str(d)!r})
import _pypolychord
out = {{}}
def ll(theta, phi):
    return float(-0.5 * np.sum((theta - 0.5) ** 2) / 0.01 - 4 * np.log(0.1 * np.sqrt(2 * np.pi)))
def prior(cube, theta):
    theta[:] = cube
def dumper(live, dead, logweights, logZ, logZerr):
    out.update(logZ=logZ, logZerr=logZerr, ndead=dead.shape[0], nlive=live.shape[0])
_pypolychord.run(ll, prior, dumper, 4, 0, 100, 8, -1, -1, False, 0, 1e-3, -1e30, -1, 0.0, False, False, False, False, False,
                 False, False, False, False, False, False, float(np.exp(-1)), True, {str(tmp_path / 'chains')!r}, "t", [1.0], [4], {{}}, 3)
print(out["logZ"], out["logZerr"], out["ndead"], out["nlive"])
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    logZ, logZerr, ndead, nlive = r.stdout.split()[-4:]
    assert abs(float(logZ)) < 5 * float(logZerr) + 0.05 and int(nlive) == 0 and int(ndead) > 500


def test_own_cpython_shim_propagates_python_exceptions(gpu, tmp_path):
    """polychordlite_b200/pypolychord/_pypolychord (csrc/pypolychord_module.cpp): an exception raised inside the Python
    likelihood unwinds through the engine's frames and comes back as that exception (_pypolychord.cpp:219-224)."""
    from polychordlite_b200 import pypolychord

    calls = {"n": 0}

    def ll(theta):
        calls["n"] += 1
        if calls["n"] > 500:
            raise KeyError("boom")
        return float(-0.5 * np.sum((theta - 0.5) ** 2) / 0.01)

    with pytest.raises(KeyError, match="boom"):
        pypolychord.run(ll, 3, nlive=50, num_repeats=6, base_dir=str(tmp_path), file_root="x", feedback=0, seed=1,
                        read_resume=False, write_resume=False)
    # the engine is usable afterwards
    out = pypolychord.run(lambda t: float(-0.5 * np.sum((t - 0.5) ** 2) / 0.01), 3, nlive=50, num_repeats=6,
                          base_dir=str(tmp_path), file_root="y", feedback=0, seed=1, read_resume=False, write_resume=False)
    assert out is not None
