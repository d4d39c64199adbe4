"""The .ini driver path, polychord_c_interface_ini (src/polychord/ini.f90 format).
The files are written here in the format of the reference's ini/*.ini (which do not travel to the GPU box)."""
import ctypes as C

import numpy as np
import pytest

from polychordlite_b200.pypolychord.output import PolyChordOutput

pytestmark = pytest.mark.gpu

HEADER = """#| an .ini file in the reference's format
[ algorithm settings ]
nlive = {nlive}
num_repeats = {R}
do_clustering = F
grade_frac = 1
precision_criterion = 0.001
[ posterior settings ]
posteriors = T
equals = F
[ output settings ]
write_resume = F
write_stats = T
write_dead = T
write_live = F
feedback = 0
seed = {seed}
base_dir = {base}
file_root = {root}
#  : name | latex name  |speed| prior type  |prior block| prior params
#--------------------------------------------------------------------
"""


def _ini_call(gpu, like_ptr, path):
    L = gpu.lib()
    L.polychord_c_interface_ini.restype = None
    L.polychord_c_interface_ini.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
    comm = C.c_int(0)
    gpu.set_option("errors_return", 1)
    try:
        L.polychord_c_interface_ini(C.cast(like_ptr, C.c_void_p), None, str(path).encode(), C.byref(comm))
    finally:
        gpu.set_option("errors_return", 0)
    return gpu.last_run_info()


def test_ini_run_with_builtin_likelihood_is_the_device_run(gpu, tmp_path):
    D = 6
    text = HEADER.format(nlive=150, R=12, seed=4, base=tmp_path, root="g6")
    text += "".join(f"P : p{i + 1}   | \\\\theta_{{{i + 1}}}  |  1  | uniform     |  1        |  0.0  1.0\n" for i in range(D))
    ini = tmp_path / "g6.ini"
    ini.write_text(text)
    info = _ini_call(gpu, gpu.lib().pc_gaussian_loglikelihood, ini)
    assert info.status == 0 and info.kernel_launches < 60          # the whole loop ran on the device (relaunches at updates only)
    ref, _ = gpu.run(gpu.make_settings(D, 0, nlive=150, num_repeats=12, seed=4))
    assert (info.ndead, info.nlike) == (ref.ndead, ref.nlike) and info.logZ == ref.logZ
    out = PolyChordOutput(str(tmp_path), "g6")
    assert abs(out.logZ - info.logZ) < 1e-12 and out.ndead == info.ndead
    assert np.allclose(out.means, 0.5, atol=0.02) and np.allclose(out.sigmas, 0.1, atol=0.02)


def test_ini_priors_and_host_likelihood(gpu, tmp_path):
    centre = np.array([0.3, 2.0, 0.7, 1.5])

    def like(theta_p, nd, phi_p, nder):
        th = np.ctypeslib.as_array(theta_p, shape=(nd,))
        phi_p[0] = float(th.sum())
        return float(-0.5 * np.sum((th - centre) ** 2) / 0.05 ** 2)
    cb = gpu.LL_CB(like)
    text = HEADER.format(nlive=80, R=8, seed=2, base=tmp_path, root="mix")
    text += "P : a | a | 1 | gaussian      | 1 | 0.0 1.0\n"
    text += "P : b | b | 1 | log_uniform   | 2 | 0.1 10.0\n"
    text += "P : c | c | 1 | uniform       | 3 | -1.0 2.0\n"
    text += "P : d | d | 1 | exponential   | 4 | 0.5\n"
    text += "D : s | \\\\Sigma\n"
    ini = tmp_path / "mix.ini"
    ini.write_text(text)
    info = _ini_call(gpu, cb, ini)
    assert info.status == 0
    out = PolyChordOutput(str(tmp_path), "mix")
    assert np.allclose(out.means[:4], centre, atol=0.03)            # a narrow likelihood: the posterior sits on its centre
    assert abs(out.means[4] - centre.sum()) < 0.06                  # the derived parameter came through
    # evidence = integral of the likelihood times the prior density at the centre (the likelihood is narrow)
    from scipy.stats import norm
    prior_density = norm.pdf(centre[0]) * (1 / (centre[1] * np.log(100))) * (1 / 3.0) * (0.5 * np.exp(-0.5 * centre[3]))
    logZ_true = np.log(prior_density) + 4 * np.log(0.05 * np.sqrt(2 * np.pi))
    assert abs(out.logZ - logZ_true) < 0.6


def test_ini_errors(gpu, tmp_path):
    info = _ini_call(gpu, gpu.lib().pc_gaussian_loglikelihood, tmp_path / "missing.ini")
    assert info.status == -6
    bad = tmp_path / "bad.ini"
    bad.write_text(HEADER.format(nlive=50, R=4, seed=1, base=tmp_path, root="bad") + "P : a | a | 1 | no_such_prior | 1 | 0 1\n")
    assert _ini_call(gpu, gpu.lib().pc_gaussian_loglikelihood, bad).status == -6


def test_ini_sorted_uniform_prior(gpu, tmp_path):
    """sorted_uniform (priors.f90:262-270): the block's parameters come out ordered; with a flat likelihood the
    posterior is the prior of three ordered uniforms on [0, 2]: means 0.5, 1.0, 1.5."""
    def like(theta_p, nd, phi_p, nder):
        th = np.ctypeslib.as_array(theta_p, shape=(nd,))
        assert th[0] <= th[1] <= th[2]
        return 0.0
    cb = gpu.LL_CB(like)
    text = HEADER.format(nlive=60, R=6, seed=3, base=tmp_path, root="srt").replace("precision_criterion = 0.001", "precision_criterion = 0.001\nmax_ndead = 600")
    text += "".join(f"P : t{i} | t_{i} | 1 | sorted_uniform | 1 | 0.0 2.0\n" for i in range(3))
    ini = tmp_path / "srt.ini"
    ini.write_text(text)
    info = _ini_call(gpu, cb, ini)
    assert info.status == 0
    out = PolyChordOutput(str(tmp_path), "srt")
    assert np.allclose(out.means, [0.5, 1.0, 1.5], atol=0.12) and abs(out.logZ) < 0.2   # Z = 1 for a flat likelihood of 1
