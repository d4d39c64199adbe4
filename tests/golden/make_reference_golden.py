"""Writes tests/golden/reference_python.json from the REFERENCE's own Python modules, imported from /root/reference
(this container only; the fixture is what travels):

  priors     every class of /root/reference/pypolychord/priors.py evaluated on fixed cube points
  settings   the attributes of /root/reference/pypolychord/settings.py's PolyChordSettings for two constructions
  stats      a <root>.stats file written by this repository's writer (pc_write_files, host-only, deterministic inputs)
             parsed by the reference's PolyChordOutput (/root/reference/pypolychord/output.py:57-99); the file's text is
             stored with the parsed fields, so the test can check both that the writer still produces these bytes and
             that the mirror parser reads them as the reference does

Run:  python tests/golden/make_reference_golden.py
"""
import importlib.util
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference/pypolychord")
sys.path.insert(0, str(ROOT))


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, REF / (name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


PRIOR_CASES = [("UniformPrior", [-2.0, 3.0]), ("GaussianPrior", [0.5, 2.0]), ("LogUniformPrior", [0.1, 10.0]),
               ("SortedUniformPrior", [-1.0, 4.0]), ("LogSortedUniformPrior", [0.01, 5.0])]


def stats_inputs():
    rng = np.random.default_rng(2024)
    D, P, ndead, nlive = 3, 2, 250, 20
    def rows(n, lo):
        logL = np.sort(lo + rng.uniform(0, 6, n))
        return np.column_stack([rng.uniform(-1, 1, (n, D)), rng.uniform(0, 1, (n, P)), logL - 0.25, logL])
    dead, live = rows(ndead, -12.0), rows(nlive, -5.0)
    logw = -np.arange(ndead) / 20.0 + dead[:, -1]
    return D, P, dead, logw, live, dict(logZ=-4.125, logZerr=0.0625, nlike=987654, num_repeats=9, seed=5)


def main():
    priors, output = _load("priors"), _load("output")
    rng = np.random.default_rng(7)
    cubes = rng.random((6, 5))
    out = {"source": "/root/reference/pypolychord (PolyChordLite 1.22.2), imported by tests/golden/make_reference_golden.py",
           "priors": {"cubes": cubes.tolist(), "cases": []}}
    for name, args in PRIOR_CASES:
        p = getattr(priors, name)(*args)
        out["priors"]["cases"].append({"class": name, "args": args, "theta": [np.asarray(p(c)).tolist() for c in cubes]})
    out["priors"]["forced_indentifiability_transform"] = [priors.forced_indentifiability_transform(c).tolist() for c in cubes]

    settings = _load("settings")
    out["settings"] = []
    for nDims, nDerived, kw in ((4, 1, {}), (7, 0, {"nlive": 50, "grade_dims": [3, 4], "grade_frac": [1.0, 2.0], "seed": 3})):
        st = settings.PolyChordSettings(nDims, nDerived, **kw)
        out["settings"].append({"nDims": nDims, "nDerived": nDerived, "kwargs": kw,
                                "attributes": {k: (float(v) if isinstance(v, np.floating) else v) for k, v in vars(st).items()}})

    from polychordlite_b200 import _capi
    D, P, dead, logw, live, kw = stats_inputs()
    with tempfile.TemporaryDirectory() as tmp:
        _capi.write_files(tmp, "gold", D, P, dead, logw, live, flags=("stats", "posteriors", "equals"), **kw)
        text = (Path(tmp) / "gold.stats").read_text()
        o = output.PolyChordOutput(tmp, "gold")
        out["stats"] = {"text": text, "parsed": {k: getattr(o, k) for k in (
            "logZ", "logZerr", "logZs", "logZerrs", "ncluster", "nposterior", "nequals", "ndead", "nlive", "nlike",
            "avnlike", "avnlikeslice")}}
    (Path(__file__).parent / "reference_python.json").write_text(json.dumps(out, indent=1))
    print("written", len(json.dumps(out)), "bytes")


if __name__ == "__main__":
    main()
