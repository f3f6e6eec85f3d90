"""The host C++ facade (slam-localization_b200/facade/localization_b200.hpp) replays the reference's own
test flows; its numbers are compared with the CPU oracle.  Without a GPU the binary must refuse to compute."""
import os
import subprocess

import numpy as np
import pytest

import parity
from slam_localization_b200 import build as slb_build
from slam_localization_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "facade_test")


def _build():
    slb_build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "facade_test.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        csrc = os.path.join(ROOT, "slam-localization_b200", "csrc")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EXE, src, "-L" + csrc, "-lslb", "-Wl,-rpath," + csrc,
                               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return EXE


def _run():
    p = subprocess.run([_build()], capture_output=True, text=True)
    out = {}
    for line in p.stdout.splitlines():
        f = line.split()
        if f:
            out.setdefault(f[0], []).append(f[1:])
    return p.returncode, out


def test_facade_compiles_and_refuses_without_gpu():
    import torch
    rc, out = _run()
    if torch.cuda.is_available():
        assert rc == 0 and "OK" in out
    else:
        assert rc == 3 and "NO_DEVICE" in out


@pytest.mark.gpu
def test_facade_replays_reference_tests(slo):
    rc, out = _run()
    assert rc == 0 and "OK" in out, out
    v = lambda k, i=0: np.array([float(x) for x in out[k][i]])
    # DATAMODEL (test/DataModelUnitTest.cpp:30-67)
    fx = synth.datamodel_fixture()
    np.testing.assert_array_equal(v("dm_plus_x"), (fx["x1"] + fx["x2"])[0])
    np.testing.assert_array_equal(v("dm_minus_x"), (fx["x1"] - fx["x2"])[0])
    np.testing.assert_array_equal(v("dm_minus_C"), (2e-10 * np.eye(3)).ravel())      # operator- adds covariances
    xr, Cr = slo.datamodel(0, fx["x1"], fx["C1"], fx["x2"], fx["C2"])
    np.testing.assert_array_equal(v("dm_fusion_x"), xr[0])
    np.testing.assert_array_equal(v("dm_fusion_C"), Cr[0].ravel())
    np.testing.assert_allclose(v("dm_fusion_x"), [0.0146635, 0.0011758085, -0.0187294], rtol=1e-9)
    xs, Cs = slo.safe_fusion(fx["x1"], fx["C1"], fx["x2"], fx["C2"])                 # :72-76
    np.testing.assert_array_equal(v("dm_safe_x"), xs[0])
    np.testing.assert_array_equal(v("dm_safe_C"), Cs[0].ravel())
    # UKFOM (test/UKFoMUnitTest.cpp:93-117)
    ux = synth.ukfom_fixture()
    mu, P, st, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU_REFBUG, slo.MM_GPS_POS, ux["mu"], ux["P"], ux["u"], ux["dt"], ux["Q"],
                                ux["z"], ux["R"])
    parity.assert_parity(slo, [0, 1, 0], v("ukfom_mu")[None], v("ukfom_sigma").reshape(1, 9, 9), mu, P)
    # USCKF_DYNAMIC (test/UsckfUnitTest.cpp:175-284)
    np.testing.assert_allclose(v("usckf_trace", 0), [0.02700075], rtol=1e-7)
    np.testing.assert_allclose(v("usckf_trace", 1), [0.03300105], rtol=1e-7)
    assert int(out["usckf_status"][0][0]) & 1          # update on the indefinite ctor-#2 covariance is flagged
    # MSCKF (test/MsckfUnitTest.cpp:151-213): two predicts
    k = 4
    mu0 = synth.identity_q(synth.STATE_BLOCKS + [0, 1] * k)[None]
    P0 = 0.025 * np.eye(12 + 6 * k)[None]
    h = 0.5 * synth.D2R
    c, s = np.cos(h), np.sin(h)
    u = np.array([[0.1, 0.1, 0.1, c ** 3 + s ** 3, s * c * c - c * s * s, c * s * c + s * c * s, c * c * s - s * s * c,
                   0.1, 0.1, 0.1, 0.1, 0.1, 0.1]])
    m, Pm = mu0, P0
    for _ in range(2):
        m, Pm, st = slo.msckf_predict(slo.PM_MSCKF_DELTAPOSE, k, m, Pm, u, 0.0, 0.01 * np.eye(12))
    parity.assert_parity(slo, synth.STATE_BLOCKS, v("msckf_mu")[None], v("msckf_P").reshape(1, 12, 12), m[:, :13],
                         parity.symmetrize_lower(Pm)[:, :12, :12])
    assert int(out["msckf_check"][0][0]) == 0           # checkSigmaPoints: Pktest == Pk, mean unmoved
    assert int(out["usckf_check"][0][0]) == 4           # ... and the LLT failure of the indefinite ctor-#2 covariance
    np.testing.assert_array_equal(v("msckf_mu_set"), v("msckf_mu") + np.r_[1.0, np.zeros(12)])
    Pset = v("msckf_P").reshape(12, 12).copy()
    Pset[np.diag_indices(12)] *= 2.0
    np.testing.assert_array_equal(v("msckf_P_set").reshape(12, 12), Pset)
    np.testing.assert_allclose(v("msckf_stats")[:2], [1.0, v("msckf_mu_set")[0]], rtol=1e-15)   # count, sum of pos.x
    # SURVEY 8f rows f2 / f4 through the facade
    state = np.zeros((1, 48))
    state[0, [6, 22, 38]] = 1.0
    state[0, 32] = 1.0
    P0 = 0.01 * np.eye(45)[None]
    F = np.eye(15)
    F[0:3, 3:6] = 0.01 * np.eye(3)
    H = np.zeros((3, 45))
    H[:, 15:18] = -np.eye(3)
    H[:, 30:33] = np.eye(3)
    err1, P1 = slo.ekf_predict(np.zeros((1, 45)), P0, F[None], 1e-4 * np.eye(15))
    P2, ret, acc = slo.ekf_update(state, P1, np.array([[1.02, 0.01, -0.01]]), H, 0.0025 * np.eye(3), gate=1)
    assert acc[0] == 1
    np.testing.assert_array_equal(v("ekf_ret"), ret[0])
    assert parity.cov_error(v("ekf_P").reshape(1, 45, 45), P2) <= parity.STEP_TOL
    velcov = np.diag([1e-2] * 3 + [1e-3] * 3)
    vel = np.array([[1, 0, 0, 0, 0, 0.1]])
    post, pcov, dpose, dcov = slo.dr_update_pose(0.01, vel, vel, velcov, np.array([[1., 2, 3, 1, 0, 0, 0]]), 1e-3 * np.eye(6)[None])
    np.testing.assert_allclose(v("dr_post"), post[0], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(v("dr_delta"), dpose[0], rtol=1e-12, atol=1e-15)
    assert parity.cov_error(v("dr_post_cov").reshape(1, 6, 6), pcov) <= parity.STEP_TOL
