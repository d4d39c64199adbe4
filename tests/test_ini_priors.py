"""The .ini driver's prior transforms (pc_ini_prior_transform, host-only) against the numpy restatement of
src/polychord/priors.f90 in oracle/priors_oracle.py: separable, sorted, adaptive sorted and
nn_adaptive_layer_gaussian families, block by block."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
import priors_oracle as po  # noqa: E402

from polychordlite_b200 import _capi  # noqa: E402

NAMES = ["", "uniform", "log_uniform", "power_uniform", "gaussian", "half_gaussian", "exponential", "sorted_uniform",
         "sorted_gaussian", "sorted_half_gaussian", "sorted_exponential", "adaptive_sorted_uniform",
         "adaptive_sorted_gaussian", "adaptive_sorted_half_gaussian", "adaptive_sorted_exponential",
         "nn_adaptive_layer_gaussian"]
PARAMS = {1: [-2.0, 3.0], 2: [0.1, 10.0], 3: [1.0, 9.0, 2.0], 4: [0.5, 2.0], 5: [1.0, 0.5], 6: [1.5]}
BASE = {7: 1, 8: 4, 9: 5, 10: 6, 11: 1, 12: 4, 13: 5, 14: 6, 15: 4}


def _ini(tmp_path, blocks):
    """blocks: list of (prior_type, size).  Returns (path, [(type, params)...], sizes)."""
    lines = ["nlive = 50", "num_repeats = 4"]
    spec, k = [], 0
    for b, (ptype, m) in enumerate(blocks, 1):
        q = PARAMS[ptype if ptype <= 6 else BASE[ptype]]
        flat = []
        for j in range(m):
            qj = [v + 0.25 * j for v in q] if ptype not in (2, 3) else q
            flat += qj
            lines.append(f"P : p{k} | p_{{{k}}} | 1 | {NAMES[ptype]} | {b} | " + " ".join(repr(v) for v in qj))
            k += 1
        spec.append((ptype, m, flat))
    path = tmp_path / "priors.ini"
    path.write_text("\n".join(lines) + "\n")
    return path, spec, k


def _transform(path, cube):
    L = _capi.lib()
    L.pc_ini_prior_transform.restype = C.c_int
    L.pc_ini_prior_transform.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    cube = np.ascontiguousarray(cube, dtype=np.float64)
    out = np.zeros_like(cube)
    rc = L.pc_ini_prior_transform(str(path).encode(), cube.ctypes.data_as(C.POINTER(C.c_double)),
                                  out.ctypes.data_as(C.POINTER(C.c_double)), len(cube))
    return rc, out


def _oracle(spec, cube):
    out, i = [], 0
    for ptype, m, flat in spec:
        out.append(po.block_htp(ptype, cube[i:i + m], flat))
        i += m
    return np.concatenate(out)


@pytest.mark.parametrize("blocks", [
    [(1, 2), (2, 1), (3, 1), (4, 2), (5, 1), (6, 2)],
    [(7, 4), (8, 3), (9, 2), (10, 3)],
    [(11, 5), (12, 4)],
    [(13, 4), (14, 5)],
    [(15, 6), (1, 1)],
    [(11, 1), (11, 2)],
])
def test_transform_matches_the_restatement(tmp_path, blocks):
    path, spec, n = _ini(tmp_path, blocks)
    rng = np.random.default_rng(17)
    for _ in range(40):
        cube = rng.random(n)
        rc, got = _transform(path, cube)
        assert rc == 0
        # tolerance: AS241 (the reference's inv_normal_cdf, utils.F90:777-966) against scipy's ndtri, ~1e-15 relative
        np.testing.assert_allclose(got, _oracle(spec, cube), rtol=1e-12, atol=1e-13)


def test_adaptive_sorted_uniform_properties(tmp_path):
    """The count coordinate covers (0.5, m - 0.5); exactly the first nfunc of the others come out ordered and the
    rest stay the plain uniform transform (priors.f90:367-385)."""
    m = 6
    path, spec, n = _ini(tmp_path, [(11, m)])
    rng = np.random.default_rng(3)
    seen = set()
    for _ in range(200):
        cube = rng.random(n)
        rc, th = _transform(path, cube)
        assert rc == 0
        assert 0.5 <= th[0] <= m - 0.5
        nfunc = int(th[0] + 0.5)
        seen.add(nfunc)
        lo = np.array([spec[0][2][2 * j] for j in range(1, m)])
        hi = np.array([spec[0][2][2 * j + 1] for j in range(1, m)])
        unit = (th[1:] - lo) / (hi - lo)
        assert np.all(np.diff(unit[:nfunc]) >= 0)
        np.testing.assert_allclose(unit[nfunc:], cube[1 + nfunc:], rtol=0, atol=1e-12)
    assert seen == set(range(1, m))


def test_nn_adaptive_layer_switches_family(tmp_path):
    path, spec, n = _ini(tmp_path, [(15, 5)])
    cube = np.array([0.1, 0.3, 0.2, 0.4, 0.6])
    rc, one = _transform(path, cube)          # one hidden layer: half-Gaussian, nothing below the mean
    assert rc == 0 and one[0] < 1.5
    mus = np.array([spec[0][2][2 * j] for j in range(2, 5)])
    assert np.all(one[2:] >= mus)
    cube[0] = 0.9
    rc, two = _transform(path, cube)          # two layers: Gaussian, cube 0.2 lies below the mean
    assert rc == 0 and two[0] >= 1.5
    assert np.any(two[2:] < mus)


def test_errors(tmp_path):
    path, _, n = _ini(tmp_path, [(1, 2)])
    assert _transform(path, np.zeros(n + 1))[0] == -7
    assert _transform(tmp_path / "missing.ini", np.zeros(2))[0] == -6
    bad = tmp_path / "bad.ini"
    bad.write_text("nlive = 5\nnum_repeats = 2\nP : a | a | 1 | no_such_prior | 1 | 0 1\n")
    assert _transform(bad, np.zeros(1))[0] == -6


def test_malformed_files_are_refused_not_crashed(tmp_path):
    """Random damage to a valid file: the parser either still yields a transform (rc 0) or reports -6/-7."""
    path, _, n = _ini(tmp_path, [(1, 2), (7, 3), (11, 4), (15, 3), (3, 1)])
    good = path.read_text().splitlines()
    rng = np.random.default_rng(23)
    junk = ["P : x", "P : a | b | c | d | e | f", "P : a | a | 1 | uniform | 1 |", "P : a | a | x | gaussian | 1 | 0 1",
            "nlive = many", "P : a | a | 1 | sorted_uniform | z | 0 1", "= = =", "P : a | a | 1 | power_uniform | 1 | 0 1",
            "grade_frac = 0.5 0.5 0.5", "P : q | q | 0 | exponential | 9 | 2.0"]
    seen = set()
    for trial in range(60):
        lines = list(good)
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, len(lines) + 1))
            if rng.random() < 0.5 and lines:
                lines.pop(min(k, len(lines) - 1))
            else:
                lines.insert(k, junk[int(rng.integers(0, len(junk)))])
        bad = tmp_path / f"bad{trial}.ini"
        bad.write_text("\n".join(lines) + "\n")
        for dims in (n, n + 1, n - 1):
            rc, out = _transform(bad, rng.random(max(dims, 1)))
            assert rc in (0, -6, -7)
            seen.add(rc)
            if rc == 0:
                assert np.all(np.isfinite(out))
    assert {-6, -7} <= seen
