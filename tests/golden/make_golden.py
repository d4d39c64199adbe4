"""Writes tests/golden/oracle_golden.json from the CPU oracle (oracle/pc_oracle.cpp).

The reference is Fortran and cannot be built in this image (no gfortran), and its own test
suite contains no numeric vectors, so the committed fixtures are (a) the analytic evidences the
reference's built-in likelihoods are normalised to (tests/golden/analytic.json, values derived in
SURVEY.md section 6) and (b) this file: outputs of the oracle on fixed seeds, so that the same
numbers can be asserted on the GPU box (where /root/reference does not exist) both for the oracle
itself and for the CUDA engine.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as O  # noqa: E402


def chain_case(D, P, R, seed, uid, like, rng):
    s = O.make_settings(D, P, nlive=10, num_repeats=R, seed=seed)
    kw = dict(prior_lo=[-5.12] * D, prior_hi=[5.12] * D) if like == "rastrigin" else {}
    cube = 0.5 + 0.05 * rng.standard_normal(D)
    rec, _ = O.calculate_points(s, cube[None, :], like=like, **kw)
    A = rng.standard_normal((D, D)) * 0.01
    chol = np.tril(A) + 0.03 * np.eye(D)
    logL = float(rec[0, -1] - 3.0)
    babies, nlike = O.slice_chain(s, rec[0], chol, logL, uid, like=like, **kw)
    return dict(D=D, P=P, R=R, seed=seed, uid=uid, like=like, seed_point=rec[0].tolist(),
                cholesky=chol.ravel().tolist(), logL=logL, nlike=int(nlike), last_baby=babies[-1].tolist(),
                babies_logL=babies[:, -1].tolist())


def run_case(D, P, nlive, R, seed, batch_K):
    s = O.make_settings(D, P, nlive=nlive, num_repeats=R, seed=seed, batch_K=batch_K)
    r, _ = O.run(s)
    return dict(D=D, P=P, nlive=nlive, R=R, seed=seed, batch_K=batch_K, ndead=int(r.ndead), nlike=int(r.nlike),
                logZ=r.logZ, logZerr=r.logZerr, nupdates=int(r.nupdates), ngenerations=int(r.ngenerations),
                nphantoms_final=int(r.nphantoms_final))


def main():
    rng = np.random.default_rng(20261017)
    L = O.lib()
    g = dict(uniforms=[], chains=[], runs=[])
    for seed, tag, uid, b in [(0, 1, 0, 0), (12345, 5, 2 ** 40 + 17, 3)]:
        g["uniforms"].append(dict(seed=seed, tag=tag, uid=uid, b=b,
                                  values=[L.oracle_uniform(seed, tag, uid, a, b) for a in range(8)]))
    g["chains"].append(chain_case(20, 2, 40, 3, 7, "gaussian", rng))
    g["chains"].append(chain_case(4, 0, 20, 1, 99, "gaussian", rng))
    g["chains"].append(chain_case(10, 0, 50, 2, 5, "rastrigin", rng))
    g["runs"].append(run_case(4, 1, 64, 8, 0, 16))
    g["runs"].append(run_case(20, 2, 200, 40, 1, 50))
    g["runs"].append(run_case(20, 2, 200, 40, 1, 0))
    (HERE / "oracle_golden.json").write_text(json.dumps(g, indent=1))
    analytic = {
        "gaussian20_unit_cube": {"logZ": -1.15e-5, "H": 17.67, "post_mean": 0.5, "post_sd": 0.1},
        "gaussian4_box_pm1": {"logZ": -2.772588722239781},
        "rastrigin2_box_5.12": {"logZ": -4.6526},
        "rastrigin10_box_5.12": {"logZ": -23.2630},
    }
    (HERE / "analytic.json").write_text(json.dumps(analytic, indent=1))


if __name__ == "__main__":
    main()
