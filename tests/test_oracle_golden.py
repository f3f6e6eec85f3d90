"""The oracle must keep reproducing the committed golden fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = dict(rtol=1e-12, atol=1e-15)


def test_ukf_golden(slo):
    g = np.load(os.path.join(G, "ukf_mtk9.npz"))
    mu, P, st, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, g["mu0"], g["P0"], g["u"], float(g["dt"]), g["Q"],
                                g["z"], g["R"])
    np.testing.assert_allclose(mu, g["mu1"], **TOL)
    np.testing.assert_allclose(P, g["P1"], **TOL)


def test_usckf_golden(slo):
    g = np.load(os.path.join(G, "usckf_n48.npz"))
    mu1, P1, _, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, 3, 9, g["mu0"], g["P0"], g["u"], float(g["dt"]),
                                   g["Q"], None, None, update=False)
    np.testing.assert_allclose(mu1, g["mu1"], **TOL)
    np.testing.assert_allclose(P1, g["P1"], **TOL)
    mu2, P2, _, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, 3, 9, mu1, P1, None, 0.0, None, g["z"], g["R"],
                                   predict=False)
    np.testing.assert_allclose(mu2, g["mu2"], **TOL)
    np.testing.assert_allclose(P2, g["P2"], **TOL)


def test_usckf_other_shapes_golden(slo):
    g = np.load(os.path.join(G, "usckf_shapes.npz"))
    for nk, nl in ((6, 6), (9, 3), (3, 0)):
        t = "_%d_%d" % (nk, nl)
        mu2, P2, st, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, nk, nl, g["mu0" + t], g["P0" + t], g["u" + t],
                                        float(g["dt"]), g["Q" + t], g["z" + t], g["R" + t])
        assert not st.any()
        np.testing.assert_allclose(mu2, g["mu2" + t], **TOL)
        np.testing.assert_allclose(P2, g["P2" + t], **TOL)
        fl, diff = slo.check_sigma_points(2, mu2, P2, nk=nk, nl=nl)
        assert not fl.any() and diff.max() < 1e-12


def test_msckf_golden(slo):
    g = np.load(os.path.join(G, "msckf_k10_f50.npz"))
    mu1, P1, _ = slo.msckf_predict(slo.PM_MSCKF_DELTAPOSE, 10, g["mu0"], g["P0"], g["u"], 0.0, g["Q"])
    np.testing.assert_allclose(mu1, g["mu_pred"], **TOL)
    np.testing.assert_allclose(P1, g["P_pred"], **TOL)
    mu2, P2, out, _, _ = slo.msckf_update(slo.MM_MSCKF_REPROJ, 10, g["mu0"], g["P0"], g["landmarks"], g["z"], g["R"])
    np.testing.assert_allclose(mu2, g["mu_upd"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(P2, g["P_upd"], rtol=1e-10, atol=1e-15)
    np.testing.assert_array_equal(out, g["outliers"])


def test_fusion_golden(slo):
    for d in (3, 6):
        g = np.load(os.path.join(G, "fusion_d%d.npz" % d))
        xo, Co = slo.datamodel(0, g["x1"], g["C1"], g["x2"], g["C2"])
        np.testing.assert_array_equal(xo, g["xo"])
        np.testing.assert_array_equal(Co, g["Co"])


def test_next_rows_golden(slo):
    g = np.load(os.path.join(G, "ekf_n45.npz"))
    err1, P1 = slo.ekf_predict(g["err0"], g["P0"], g["F"], g["Q"])
    np.testing.assert_allclose(err1, g["err1"], **TOL)
    np.testing.assert_allclose(P1, g["P1"], **TOL)
    P2, ret, acc = slo.ekf_update(g["mu0"], P1, g["z"], g["H"], g["R"], gate=1)
    np.testing.assert_allclose(P2, g["P2"], **TOL)
    np.testing.assert_allclose(ret, g["ret"], **TOL)
    np.testing.assert_array_equal(acc, g["acc"])
    assert 0 < acc.sum() < len(acc)
    mu3, P3, acc3 = slo.ekf_single_update(g["mu0"], err1, P2, g["zs"], g["Hs"], g["R"], gate=1)
    np.testing.assert_allclose(mu3, g["mu3"], **TOL)
    np.testing.assert_allclose(P3, g["P3"], **TOL)
    np.testing.assert_array_equal(acc3, g["acc3"])
    g = np.load(os.path.join(G, "safe_fusion_d3.npz"))
    xo, Co = slo.safe_fusion(g["x1"], g["C1"], g["x2"], g["C2"])
    np.testing.assert_array_equal(xo, g["xo"])
    np.testing.assert_array_equal(Co, g["Co"])
    g = np.load(os.path.join(G, "deadreckon.npz"))
    out = slo.dr_update_pose(float(g["dt"]), g["vel0"], g["vel1"], g["velcov"], g["prev_pose"], g["prev_cov"])
    for a, k in zip(out, ("post", "pcov", "dpose", "dcov")):
        np.testing.assert_allclose(a, g[k], **TOL)
