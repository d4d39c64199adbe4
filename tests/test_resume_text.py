"""The reference's TEXT resume layout (read_write.F90:219-288 writer, :384-476 reader) on
the host: the fixture tests/golden/reference_cube_samples.resume was written by the reference's own Python writer
(pypolychord/polychord.py:650-789 _make_resume_file, driven by tests/golden/make_reference_resume.py); the engine's reader
must take it apart as the Fortran reader does, and the engine's writer must give the same bytes back."""
import json
from pathlib import Path

import numpy as np
import pytest

from polychordlite_b200 import _capi as capi

GOLD = Path(__file__).parent / "golden"


def test_reader_takes_the_reference_written_file_apart():
    meta = json.loads((GOLD / "reference_cube_samples.json").read_text())
    r = capi.resume_text_probe(GOLD / "reference_cube_samples.resume")
    assert (r["nDims"], r["nDerived"], r["ndead"], r["ncluster"], r["ncluster_dead"]) == (meta["nDims"], meta["nDerived"], 0, 1, 0)
    assert (r["nlive"], r["nphantom"], r["nlike"]) == (meta["nlive"], 0, meta["nlive"])
    assert r["logZ"] == -1e30 and r["logZ2"] == -1e30 and r["logX"] == 0.0 and r["logX_last_update"] == 0.0
    cubes = np.array(meta["cubes"])
    logL = -0.5 * ((cubes - meta["mu"]) ** 2).sum(axis=1) / meta["sigma"] ** 2 - meta["nDims"] * np.log(meta["sigma"] * np.sqrt(2 * np.pi))
    assert abs(r["logL_min"] - logL.min()) < 1e-12 and abs(r["logL_max"] - logL.max()) < 1e-12


def test_writer_gives_the_reference_written_bytes_back(tmp_path):
    out = tmp_path / "again.resume"
    capi.resume_text_probe(GOLD / "reference_cube_samples.resume", out)
    assert out.read_bytes() == (GOLD / "reference_cube_samples.resume").read_bytes()


@pytest.mark.parametrize("damage", ["truncate", "letters", "count", "header"])
def test_damaged_files_are_refused(tmp_path, damage):
    text = (GOLD / "reference_cube_samples.resume").read_text().splitlines()
    if damage == "truncate":
        text = text[:40]
    elif damage == "letters":
        text[1] = "        four"
    elif damage == "count":
        i = text.index("=== live points ===")
        text[i + 2] = text[i + 2][:24 * 5]            # a live point with too few columns
    else:
        text[2] = "Number of derived parameters"      # the '===' of a section header is missing
    bad = tmp_path / "bad.resume"
    bad.write_text("\n".join(text) + "\n")
    with pytest.raises(ValueError):
        capi.resume_text_probe(bad)
