"""Resume (replaces read_write.F90:219-476): a run interrupted after a few updates and
picked up from its resume file ends BIT-IDENTICAL to the uninterrupted run (every random number is counter-addressed and
the file carries the run's seed), for device likelihoods, with clustering, and for host callbacks; a finished run's file
returns the finished result; a file of a different run shape is a fatal error (read_write.F90:402-417)."""
import numpy as np
import pytest

from polychordlite_b200 import pypolychord
from polychordlite_b200.pypolychord.priors import UniformPrior

pytestmark = pytest.mark.gpu


class Stop(Exception):
    pass


def _final(dumps):
    return dumps[-1]["dead"], dumps[-1]["logweights"], dumps[-1]["logZ"]


@pytest.mark.parametrize("clustering,like,D,box", [(False, "gaussian", 6, None), (True, "rastrigin", 2, 5.12)])
def test_interrupted_run_resumes_bit_identical(gpu, tmp_path, clustering, like, D, box):
    kw = dict(prior_lo=[-box] * D, prior_hi=[box] * D) if box else {}
    st = gpu.make_settings(D, 0, nlive=200, num_repeats=2 * D, seed=11, do_clustering=clustering)
    ref, ref_dumps = gpu.run(st, like=like, want_dump=True, **kw)            # uninterrupted, no resume file involved
    path = tmp_path / "r.resume"
    gpu.set_option("errors_return", 1)
    try:
        # first leg: stopped from inside the 4th dumper call, the resume file rewritten at every update
        gpu.set_resume(path, write=True)
        gpu.set_option("resume_interval", 0.0)
        with pytest.raises(RuntimeError):
            gpu.run(st, like=like, abort_after_dumps=4, **kw)
        assert path.exists()
        # second leg: pick the run up
        gpu.set_resume(path, read=True)
        res, res_dumps = gpu.run(st, like=like, want_dump=True, **kw)
    finally:
        gpu.set_resume()
        gpu.set_option("resume_interval", 1.0)
        gpu.set_option("errors_return", 0)
    assert (res.ndead, res.nlike, res.nupdates) == (ref.ndead, ref.nlike, ref.nupdates)
    assert res.logZ == ref.logZ
    d0, w0, z0 = _final(ref_dumps)
    d1, w1, z1 = _final(res_dumps)
    assert np.array_equal(d0, d1) and np.array_equal(w0, w1) and z0 == z1


def test_finished_run_and_shape_mismatch(gpu, tmp_path):
    path = tmp_path / "f.resume"
    st = gpu.make_settings(4, 1, nlive=100, num_repeats=8, seed=3)
    gpu.set_option("errors_return", 1)
    try:
        gpu.set_resume(path, write=True)
        a, da = gpu.run(st, want_dump=True)
        gpu.set_resume(path, read=True)
        b, db = gpu.run(st, want_dump=True)                                  # nothing left to sample
        assert (a.ndead, a.nlike, a.logZ) == (b.ndead, b.nlike, b.logZ) and b.kernel_launches == 0
        assert np.array_equal(da[-1]["dead"], db[-1]["dead"])
        with pytest.raises(RuntimeError):                                    # another run's file
            gpu.run(gpu.make_settings(4, 1, nlive=120, num_repeats=8, seed=3))
    finally:
        gpu.set_resume()
        gpu.set_option("errors_return", 0)


def test_python_api_resumes_a_host_callback_run(gpu, tmp_path):
    """write_resume / read_resume through pypolychord.run with a Python likelihood that fails half-way."""
    def like(theta):
        return -float(np.sum(np.asarray(theta) ** 2)) / 0.02

    n = [0]

    def flaky(theta):
        n[0] += 1
        if n[0] == 6000:
            raise Stop()
        return like(theta)

    kw = dict(prior=UniformPrior(-1, 1), nlive=60, num_repeats=6, feedback=0, do_clustering=False, base_dir=str(tmp_path),
              seed=5, write_stats=False, write_live=False, write_dead=False, write_prior=False, posteriors=False, equals=False)
    ref = pypolychord.run(like, 3, file_root="ref", write_resume=False, read_resume=False, **kw)
    with pytest.raises(Stop):
        pypolychord.run(flaky, 3, file_root="leg", write_resume=True, read_resume=False, **kw)
    assert (tmp_path / "leg.resume").exists()
    res = pypolychord.run(like, 3, file_root="leg", write_resume=True, read_resume=True, **kw)
    assert res.equals(ref) and res.logZ == ref.logZ



def test_damaged_resume_file_is_refused(gpu, tmp_path):
    """A file that passes the shape header but whose counters do not fit its arrays must not reach the kernel."""
    import struct
    st = gpu.make_settings(6, 0, nlive=200, num_repeats=12, seed=11)
    path = tmp_path / "r.resume"
    gpu.set_option("errors_return", 1)
    try:
        gpu.set_resume(path, write=True)
        gpu.set_option("resume_interval", 0.0)
        with pytest.raises(RuntimeError):
            gpu.run(st, abort_after_dumps=3)
        raw = bytearray(path.read_bytes())
        header = 136                                       # sizeof(ResumeHeader) (csrc/pc_engine.cu: 124 bytes of ints, padded, one double)
        ndead_at = header + 8 * 8                          # DevRun: eight doubles, then ndead
        (ndead,) = struct.unpack_from("<q", raw, ndead_at)
        assert 0 < ndead < 10 ** 6
        struct.pack_into("<q", raw, ndead_at, ndead + 12345)
        path.write_bytes(bytes(raw))
        gpu.set_resume(path, read=True)
        with pytest.raises(RuntimeError):
            gpu.run(st)
    finally:
        gpu.set_resume()
        gpu.set_option("resume_interval", 1.0)
        gpu.set_option("errors_return", 0)
