"""CPU checks of the oracle for the SURVEY 8(f) "next" rows (oracle/slo_next.hpp): f2 error-state EKF with
Joseph-form update, f3 safeFusion (restated Eigen JacobiSVD), f4 dead reckoning / transform composition with
uncertainty.  The reference asserts none of these outputs (parity unpinned), so the restatement is pinned by
independent numpy formulations, analytic properties and finite differences."""
import numpy as np

from slam_localization_b200 import synth


def test_jacobi_svd_restatement(slo):
    rng = np.random.default_rng(0)
    for n in (2, 3, 6, 10):
        A = rng.normal(size=(n, n))
        U, sv = slo.jacobi_svd(A)
        np.testing.assert_allclose(sv, np.linalg.svd(A)[1], rtol=1e-12)
        np.testing.assert_allclose(U.T @ U, np.eye(n), atol=1e-14)
        S = A @ A.T + 0.1 * np.eye(n)                     # symmetric PSD: U diagonalises it
        U, sv = slo.jacobi_svd(S)
        np.testing.assert_allclose(U.T @ S @ U, np.diag(sv), atol=1e-12 * sv[0])
        assert np.all(np.diff(sv) <= 0)                  # sorted descending


def _safe_np(slo, x1, C1, x2, C2):
    I1, I2 = np.linalg.inv(C1), np.linalg.inv(C2)
    U1, s1 = slo.jacobi_svd(I1)
    sq = np.diag(np.sqrt(s1))
    isq = np.linalg.inv(sq)
    U2, s2 = slo.jacobi_svd(isq @ U1.T @ I2 @ U1 @ isq)
    T = U2.T @ sq @ U1                                    # DataModel.hpp:106 as written
    res = np.where(s2 < 1.0, T @ x1, T @ x2)
    D3 = np.diag(np.where(s2 < 1.0, 1.0, s2))
    Ti = np.linalg.inv(T)
    return Ti @ res, Ti @ np.linalg.inv(D3) @ Ti.T


def test_safe_fusion_against_numpy(slo):
    sc = synth.safe_fusion_scenario(64, log_spread=1.0)
    xo, Co = slo.safe_fusion(sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    for i in range(64):
        xr, Cr = _safe_np(slo, sc["x1"][i], sc["C1"][i], sc["x2"][i], sc["C2"][i])
        np.testing.assert_allclose(xo[i], xr, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(Co[i], Cr, rtol=1e-7, atol=1e-12)


def test_safe_fusion_reference_inputs(slo):
    fx = synth.safe_fusion_fixture()                       # test/DataModelUnitTest.cpp:66-74
    xo, Co = slo.safe_fusion(fx["x1"], fx["C1"], fx["x2"], fx["C2"])
    # equal covariances: I1 = 1e10 I, so I2' = I exactly and every D2(i,i) = 1 is not < 1: the result takes
    # data2 through T and back, with covariance T^-1 T^-T = C1
    np.testing.assert_allclose(xo, fx["x2"], rtol=1e-12)
    np.testing.assert_allclose(Co, fx["C1"], rtol=1e-12, atol=1e-26)


def test_ekf_against_numpy(slo):
    sc = synth.ekf_scenario(16, seed=3)
    err, P = slo.ekf_predict(sc["err"], sc["P"], sc["F"], sc["Q"])
    o = 30
    for i in range(16):
        F, P0 = sc["F"][i], sc["P"][i]
        Pn = P0.copy()
        Pn[o:, o:] = F @ P0[o:, o:] @ F.T + sc["Q"]
        for b in (0, 15):
            Pn[b:b + 15, o:] = P0[b:b + 15, o:] @ F.T
            Pn[o:, b:b + 15] = F @ P0[o:, b:b + 15]
        np.testing.assert_allclose(P[i], Pn, rtol=1e-12, atol=1e-16)
        np.testing.assert_allclose(err[i][o:], F @ sc["err"][i][o:], rtol=1e-12)
        np.testing.assert_array_equal(err[i][:o], sc["err"][i][:o])
    P2, ret, acc = slo.ekf_update(sc["mu"], sc["P"], sc["z"], sc["H"], sc["R"], gate=0)
    assert acc.all() and not ret.any()
    H, R = sc["H"], sc["R"]
    for i in range(16):
        P0 = sc["P"][i]
        S = H @ P0 @ H.T + R
        K = P0 @ H.T @ np.linalg.inv(S)
        IKH = np.eye(45) - K @ H
        Pj = IKH @ P0 @ IKH.T + K @ R @ K.T
        np.testing.assert_allclose(P2[i], 0.5 * (Pj + Pj.T), rtol=1e-10, atol=1e-15)
        np.testing.assert_array_equal(P2[i], P2[i].T)           # :359 guarantees symmetry
        assert np.linalg.eigvalsh(P2[i]).min() > 0
    # gate: dof = m - 1 = 2 -> 5.99 (UsckfError.hpp:350)
    zbad = sc["z"] + 10.0
    P3, ret, acc = slo.ekf_update(sc["mu"], sc["P"], zbad, H, R, gate=1)
    assert not acc.any()
    np.testing.assert_array_equal(P3, sc["P"])
    np.testing.assert_allclose(ret, zbad - synth.ekf_vectorize(sc["mu"]) @ H.T, rtol=1e-12)


def test_ekf_single_update_and_clone(slo):
    sc = synth.ekf_scenario(8, seed=4)
    mu, P, acc = slo.ekf_single_update(sc["mu"], sc["err"], sc["P"], sc["zs"], sc["Hs"], sc["R"], gate=0)
    o = 30
    for i in range(8):
        Pk = sc["P"][i][o:, o:]
        Hs, R = sc["Hs"], sc["R"]
        S = Hs @ Pk @ Hs.T + R
        K = Pk @ Hs.T @ np.linalg.inv(S)
        xk = sc["err"][i][o:] + K @ (sc["zs"][i] - Hs @ sc["err"][i][o:])
        s0, s1 = sc["mu"][i][32:], mu[i][32:]
        np.testing.assert_allclose(s1[0:6], s0[0:6] + xk[0:6], rtol=1e-12)
        np.testing.assert_allclose(s1[10:16], s0[10:16] + xk[9:15], rtol=1e-12)
        assert abs(np.linalg.norm(s1[6:10]) - 1.0) < 1e-15
        np.testing.assert_array_equal(mu[i][:32], sc["mu"][i][:32])
        # only the statek_i block of the covariance moves
        Pd = P[i] - sc["P"][i]
        Pd[o:, o:] = 0
        assert not Pd.any()
    mu, err, P = slo.ekf_clone(sc["mu"], sc["err"], sc["P"])
    for a in range(3):
        np.testing.assert_array_equal(mu[:, 16 * a:16 * a + 16], sc["mu"][:, 32:48])
        for b in range(3):
            np.testing.assert_array_equal(P[:, 15 * a:15 * a + 15, 15 * b:15 * b + 15], sc["P"][:, 30:, 30:])


def _r2q(r):
    th = np.linalg.norm(r)
    return np.array([1.0, 0, 0, 0]) if th < 1e-12 else np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * r / th])


def _q2r(q):
    q = q / np.linalg.norm(q)
    q = -q if q[0] < 0 else q
    n = np.linalg.norm(q[1:])
    return np.zeros(3) if n == 0 else 2 * np.arctan2(n, q[0]) * q[1:] / n


def _qmul(a, b):
    return np.concatenate([[a[0] * b[0] - a[1:] @ b[1:]], a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:])])


def _rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_transform_composition_matches_finite_differences(slo):
    """The Pennec/Thirion Jacobians of Transform.cpp:65-138 are small-angle series: at 0.05 rad they agree with
    the numerically differentiated composition to a few 1e-3, which catches any transcription slip."""
    rng = np.random.default_rng(1)

    def comp(p2, p1):
        q2, q1 = _r2q(p2[:3]), _r2q(p1[:3])
        return np.concatenate([_q2r(_qmul(q2, q1)), _rot(q2) @ p1[3:] + p2[3:]])

    for _ in range(5):
        p2 = np.concatenate([rng.normal(size=3) * 0.05, rng.normal(size=3)])
        p1 = np.concatenate([rng.normal(size=3) * 0.05, rng.normal(size=3)])
        J1, J2 = np.zeros((6, 6)), np.zeros((6, 6))
        for k in range(6):
            d = np.zeros(6)
            d[k] = 1e-6
            J1[:, k] = (comp(p2, p1 + d) - comp(p2, p1 - d)) / 2e-6
            J2[:, k] = (comp(p2 + d, p1) - comp(p2 - d, p1)) / 2e-6
        A = rng.normal(size=(6, 6))
        c1 = A @ A.T * 1e-3
        A = rng.normal(size=(6, 6))
        c2 = A @ A.T * 1e-3
        pose2 = np.concatenate([p2[3:], _r2q(p2[:3])])[None]
        pose1 = np.concatenate([p1[3:], _r2q(p1[:3])])[None]
        po, co = slo.transform_compose(pose2, c2[None], pose1, c1[None])
        ref = J1 @ c1 @ J1.T + J2 @ c2 @ J2.T
        assert np.abs(co[0] - ref).max() / np.abs(ref).max() < 5e-3
        pc = comp(p2, p1)
        np.testing.assert_allclose(po[0][:3], pc[3:], atol=1e-14)
        np.testing.assert_allclose(_q2r(po[0][3:]), pc[:3], atol=1e-14)


def test_dead_reckon_update_pose(slo):
    sc = synth.deadreckon_scenario(32)
    post, pcov, dpose, dcov = slo.dr_update_pose(sc["dt"], sc["vel0"], sc["vel1"], sc["velcov"], sc["prev_pose"], sc["prev_cov"])
    dt = sc["dt"]
    for i in range(32):
        # DeadReckon.hpp:45,48-52: translation = dt/2 (v0 + v1); cov blocks swap (orientation first) and scale by dt^2
        np.testing.assert_allclose(dpose[i][:3], dt / 2 * (sc["vel0"][i][:3] + sc["vel1"][i][:3]), rtol=1e-14)
        np.testing.assert_allclose(dcov[i][:3, :3], sc["velcov"][3:, 3:] * dt * dt, rtol=1e-12)
        np.testing.assert_allclose(dcov[i][3:, 3:], sc["velcov"][:3, :3] * dt * dt, rtol=1e-12)
        assert abs(np.linalg.norm(dpose[i][3:]) - 1) < 1e-15
        # :270-272 is a third-order integration of the angular rate whose first-order term is
        # (0.75 Omega(w0) - 0.25 Omega(w1)) dt with Omega carrying no 1/2 (:255-263): as written, the quaternion's
        # vector part is (0.75 w0 - 0.25 w1) dt, i.e. twice the usual rotation angle
        w = 0.75 * sc["vel0"][i][3:] - 0.25 * sc["vel1"][i][3:]
        np.testing.assert_allclose(_q2r(dpose[i][3:]), 2.0 * w * dt, atol=5e-6)
        # postPose = prevPose * deltaPose
        q = _qmul(sc["prev_pose"][i][3:], dpose[i][3:])
        np.testing.assert_allclose(_rot(post[i][3:]), _rot(q), atol=1e-13)
        np.testing.assert_allclose(post[i][:3], _rot(sc["prev_pose"][i][3:]) @ dpose[i][:3] + sc["prev_pose"][i][:3], atol=1e-13)
        np.testing.assert_allclose(pcov[i], pcov[i].T, atol=1e-18)
        assert np.linalg.eigvalsh(0.5 * (pcov[i] + pcov[i].T)).min() > 0
    # results are per-instance: threads do not change them
    post2, pcov2, _, _ = slo.dr_update_pose(sc["dt"], sc["vel0"], sc["vel1"], sc["velcov"], sc["prev_pose"], sc["prev_cov"], nthreads=4)
    np.testing.assert_array_equal(post, post2)
    np.testing.assert_array_equal(pcov, pcov2)


# ---- f1: Msckf::update, EKF flavour with QR compression (Msckf.hpp:297-349,756-816) -------------------------------
MS_BLOCKS = [0, 1, 0, 0] + [0, 1] * 10


def test_householder_qr_restatement(slo):
    rng = np.random.default_rng(0)
    A = rng.normal(size=(40, 12))
    A[:, 3] = 0.0                                         # a zero column: tau = 0, R(3,3) = 0 (Eigen's convention)
    QR, tau, Q = slo.householder_qr(A)
    R = np.triu(QR)[:12]
    np.testing.assert_allclose(Q @ R, A, atol=1e-13)
    np.testing.assert_allclose(Q.T @ Q, np.eye(12), atol=1e-14)
    assert tau[3] != 0.0 or R[3, 3] == 0.0
    # sign convention: beta = -sign(c0) |x|
    QR2, tau2, _ = slo.householder_qr(np.array([[3.0], [4.0]]))
    assert QR2[0, 0] == -5.0 and abs(tau2[0] - 1.6) < 1e-15


def test_reproj_jacobian_matches_finite_differences(slo):
    sc = synth.msckf_scenario(2, seed=1, k=10, nfeat=50)
    z, H = slo.msckf_reproj_jac(10, sc["mu"][0], sc["landmarks"])
    Hn = np.zeros_like(H)
    for c in range(72):
        d = np.zeros(72)
        d[c] = 1e-6
        zp, _ = slo.msckf_reproj_jac(10, slo.boxplus(MS_BLOCKS, sc["mu"][0], d), sc["landmarks"])
        zm, _ = slo.msckf_reproj_jac(10, slo.boxplus(MS_BLOCKS, sc["mu"][0], -d), sc["landmarks"])
        Hn[:, c] = (zp - zm) / 2e-6
    assert np.abs(H - Hn).max() < 1e-8
    assert not H[:, :12].any()                          # statek is not observed by this model


def test_msckf_ekf_update_equals_plain_kalman_update_without_outliers(slo):
    """With R = sigma^2 I the QR compression is lossless: the update equals K = P H^T (H P H^T + R)^-1 on the
    uncompressed rows (the discarded rows are orthogonal to range(H) and uncorrelated with the kept ones)."""
    sc = synth.msckf_scenario(6, seed=3, k=10, nfeat=50)
    mu, P, out, st = slo.msckf_update_ekf(slo.MM_MSCKF_REPROJ, 10, sc["mu"], sc["P"], sc["landmarks"], sc["z"], sc["R"], gate=False)
    assert not out.any() and not st.any()
    for i in range(6):
        z, H = slo.msckf_reproj_jac(10, sc["mu"][i], sc["landmarks"])
        P0 = sc["P"][i]
        S = H @ P0 @ H.T + sc["R"]
        K = P0 @ H.T @ np.linalg.inv(S)
        np.testing.assert_allclose(P[i], P0 - K @ S @ K.T, rtol=1e-8, atol=1e-14)
        np.testing.assert_allclose(mu[i], slo.boxplus(MS_BLOCKS, sc["mu"][i], K @ (sc["z"][i] - z)), rtol=1e-9, atol=1e-12)


def test_msckf_ekf_update_outliers_and_row_shortage(slo):
    sc = synth.msckf_scenario(16, seed=2, k=10, nfeat=50, outlier_frac=0.1)
    mu, P, out, st = slo.msckf_update_ekf(slo.MM_MSCKF_REPROJ, 10, sc["mu"], sc["P"], sc["landmarks"], sc["z"], sc["R"], gate=True)
    assert out.max() >= 15 and (st == 16).any() and (st == 0).any()
    # fewer than DOF rows left: reduceDimension cannot run (Msckf.hpp:808) -> flagged and left unchanged
    short = st == 16
    assert np.all(2 * (50 - out[short]) + out[short] * 0 < 72 + 2 * out[short])   # each outlier removes two rows
    np.testing.assert_array_equal(mu[short], sc["mu"][short])
    np.testing.assert_array_equal(P[short], sc["P"][short])
    ok = st == 0
    assert np.linalg.eigvalsh(P[ok]).min() > 0
    # threads do not change results
    mu2, P2, out2, st2 = slo.msckf_update_ekf(slo.MM_MSCKF_REPROJ, 10, sc["mu"], sc["P"], sc["landmarks"], sc["z"], sc["R"],
                                              gate=True, nthreads=4)
    np.testing.assert_array_equal(mu, mu2)
    np.testing.assert_array_equal(out, out2)
