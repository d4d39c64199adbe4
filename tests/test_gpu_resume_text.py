"""Runs continued from resume files in the reference's TEXT layout (read_write.F90:219-476):
the file pypolychord's _make_resume_file wrote for cube_samples (tests/golden/reference_cube_samples.resume) starts the
same run the oracle does from the same live points; a file this engine wrote with the option "resume_text" is read back
and the run goes on to a sound evidence; a file of another problem is fatal (read_write.F90:402-417)."""
import json
import shutil
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def test_run_from_the_reference_written_file_matches_the_oracle(gpu, oracle, tmp_path):
    meta = json.loads((GOLD / "reference_cube_samples.json").read_text())
    D, P, N, R, K = meta["nDims"], meta["nDerived"], meta["nlive"], meta["num_repeats"], 6
    path = tmp_path / "cs.resume"
    shutil.copy(GOLD / "reference_cube_samples.resume", path)
    gpu.set_option("batch_K", K)
    gpu.set_resume(path, read=True)
    try:
        info, _ = gpu.run(gpu.make_settings(D, P, nlive=N, num_repeats=R, seed=5))
    finally:
        gpu.set_resume()
        gpu.set_option("batch_K", 0)
    oracle.set_initial_cubes(np.array(meta["cubes"]))
    oi, _ = oracle.run(oracle.make_settings(D, P, nlive=N, num_repeats=R, seed=5, batch_K=K))
    assert (info.ndead, info.nlike, info.nupdates) == (oi.ndead, oi.nlike, oi.nupdates)
    assert abs(info.logZ - oi.logZ) < 1e-7


def test_text_export_is_read_back_and_the_run_goes_on(gpu, tmp_path):
    D = 6
    st = gpu.make_settings(D, 0, nlive=200, num_repeats=2 * D, seed=11)
    ref, _ = gpu.run(st)
    path = tmp_path / "t.resume"
    gpu.set_option("errors_return", 1)
    gpu.set_option("resume_text", 1)
    gpu.set_option("resume_interval", 0.0)
    try:
        gpu.set_resume(path, write=True)
        with pytest.raises(RuntimeError):
            gpu.run(st, abort_after_dumps=4)
        first = path.read_text().splitlines()[0]
        assert first == "=== Number of dimensions ==="
        leg = gpu.resume_text_probe(path)
        assert (leg["nDims"], leg["nDerived"], leg["ncluster"], leg["nlive"]) == (D, 0, 1, 200)
        assert 0 < leg["ndead"] < ref.ndead and leg["nphantom"] > 0 and np.isfinite(leg["logZ"]) and leg["logX"] < 0.0
        gpu.set_resume(path, read=True)
        res, _ = gpu.run(st)
        with pytest.raises(RuntimeError):   # resume error: nDims does not match
            gpu.run(gpu.make_settings(D + 1, 0, nlive=200, num_repeats=2 * D, seed=11))
    finally:
        gpu.set_resume()
        gpu.set_option("resume_interval", 1.0)
        gpu.set_option("resume_text", 0)
        gpu.set_option("errors_return", 0)
    # the continuation is a run of its own (the text layout carries no chain counter): same problem, same evidence
    assert abs(res.logZ - ref.logZ) < 4 * np.hypot(res.logZerr, ref.logZerr)
    assert abs(res.logZ) < 5 * res.logZerr          # the Gaussian is normalised: log Z = 0
    assert 0.7 * ref.ndead < res.ndead < 1.3 * ref.ndead and res.ndead > leg["ndead"]
