"""Analytic known answers for the oracle (SURVEY.md 4.4 iii/iv) plus the reference's own test inputs
replayed through it (test/UsckfUnitTest.cpp:175-284, test/UKFoMUnitTest.cpp:93-117,
test/DataModelUnitTest.cpp:30-67).  The reference asserts none of these outputs; the analytic cases
are what pins them."""
import numpy as np

from slam_localization_b200 import synth


def test_datamodel_fixture_analytic(slo):
    fx = synth.datamodel_fixture()
    xo, Co = slo.datamodel(0, fx["x1"], fx["C1"], fx["x2"], fx["C2"])
    np.testing.assert_allclose(xo, fx["x_expected"], rtol=1e-12)
    np.testing.assert_allclose(xo[0], [0.0146635, 0.0011758085, -0.0187294], rtol=1e-9)
    np.testing.assert_allclose(Co, fx["C_expected"], rtol=1e-12, atol=1e-26)


def test_datamodel_default_and_addsub(slo):
    x, Cv = slo.datamodel_default(3)
    assert np.all(x == 0) and np.array_equal(Cv, 1e-10 * np.eye(3))      # DataModel.hpp:32-36
    sc = synth.fusion_scenario(4, d=3)
    xp, Cp = slo.datamodel(+1, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    xm, Cm = slo.datamodel(-1, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    np.testing.assert_array_equal(xp, sc["x1"] + sc["x2"])
    np.testing.assert_array_equal(xm, sc["x1"] - sc["x2"])
    np.testing.assert_array_equal(Cp, sc["C1"] + sc["C2"])
    np.testing.assert_array_equal(Cm, sc["C1"] + sc["C2"])               # operator- ADDS covariances (:149)


def test_ukf_identity_process_adds_Q(slo):
    sc = synth.ukfom_scenario(8, seed=21, p_scale=1e-4)
    u0 = np.zeros_like(sc["u"])
    # dt = 0: g is the identity map, so sigma-point regeneration must return (mu, P + Q)
    mu, P, st, it = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, sc["mu"], sc["P"], u0, 0.0, sc["Q"], None, None,
                                 update=False)
    for i in range(8):
        assert np.max(np.abs(slo.boxminus([0, 1, 0], mu[i], sc["mu"][i]))) < 1e-12
    np.testing.assert_allclose(P, sc["P"] + sc["Q"], rtol=1e-9, atol=1e-16)
    assert not st.any()


def test_ukf_linear_gps_update_is_kalman(slo):
    sc = synth.ukfom_scenario(8, seed=22, p_scale=1e-8, cond=10.0, r_sigma=1e-4)
    mu, P, st, it = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, sc["mu"], sc["P"], None, 0.0, None, sc["z"],
                                 sc["R"], predict=False)
    H = np.zeros((3, 9))
    H[:, :3] = np.eye(3)
    for i in range(8):
        S = H @ sc["P"][i] @ H.T + sc["R"]
        K = sc["P"][i] @ H.T @ np.linalg.inv(S)
        P_kf = sc["P"][i] - K @ S @ K.T
        # exact for the vector blocks; the SO3 block re-estimation (apply_delta) is exact only to
        # second order in the ~1e-4 rad sigma-point spread
        vv = np.r_[0:3, 6:9]
        np.testing.assert_allclose(P[i][np.ix_(vv, vv)], P_kf[np.ix_(vv, vv)], rtol=1e-9, atol=1e-22)
        np.testing.assert_allclose(P[i], P_kf, rtol=0, atol=1e-3 * np.abs(P_kf).max())
        d_kf = K @ (sc["z"][i] - sc["mu"][i][:3])
        np.testing.assert_allclose(slo.boxminus([0, 1, 0], mu[i], sc["mu"][i]), d_kf, rtol=1e-3, atol=1e-10)


def test_ukfom_reference_fixture_runs_clean(slo):
    fx = synth.ukfom_fixture()
    mu, P, st, it = slo.ukf_step(9, slo.PM_UKFOM_IMU_REFBUG, slo.MM_GPS_POS, fx["mu"], fx["P"], fx["u"], fx["dt"],
                                 fx["Q"], fx["z"], fx["R"])
    assert st[0] == 0
    # GPS (1,0,0) with R = 1e-8 against a 1e-3 prior: the position snaps to the measurement
    np.testing.assert_allclose(mu[0, :3], [1.0, 0, 0], atol=2e-5)
    assert np.allclose(P[0], P[0].T, atol=1e-18) and np.linalg.eigvalsh(P[0]).min() > 0


def _usckf_dynamic(slo):
    fx = synth.usckf_unit_test_fixture()
    single = synth.identity_q(synth.STATE_BLOCKS)[None]
    mu, P = slo.usckf_ctor_single(single, fx["P0_single"][None])          # UsckfUnitTest.cpp:193
    mu, P = slo.usckf_set_measurement(slo.STATEK, 0, 0, mu, P, fx["featuresVO"][None], fx["featuresVOCov"])
    mu, P = slo.usckf_set_measurement(slo.STATEK_L, 3, 0, mu, P, fx["featuresICP"][None], fx["featuresICPCov"])
    mu, P = slo.usckf_set_measurement(slo.STATEK, 3, 9, mu, P, fx["featuresVO2"][None], fx["featuresVO2Cov"])
    return fx, mu, P


def test_usckf_ctor2_block_pattern_is_indefinite(slo):                    # quirk Q13
    fx, mu, P = _usckf_dynamic(slo)
    P0 = fx["P0_single"]
    Z = np.zeros((12, 12))
    np.testing.assert_array_equal(P[0][:36, :36], np.block([[P0, P0, Z], [P0, P0, P0], [Z, P0, P0]]))
    assert np.linalg.eigvalsh(P[0][:36, :36]).min() < 0
    np.testing.assert_array_equal(P[0][36:39, 36:39], fx["featuresVO2Cov"])
    np.testing.assert_array_equal(P[0][39:, 39:], fx["featuresICPCov"])
    np.testing.assert_array_equal(mu[0][39:42], fx["featuresVO2"])


def test_usckf_unit_test_predict_traces(slo):
    """trace(P_ii) after the two predicts of USCKF_DYNAMIC; 0.02700075 / 0.03300105 are the values the
    survey observed with its own throw-away numpy restatement."""
    fx, mu, P = _usckf_dynamic(slo)
    u = np.r_[fx["velo"], fx["angvelo"]][None]
    Q = synth.usckf_process_noise(fx["dt"])
    want = [0.02700075, 0.03300105]
    for step in range(2):
        mu, P, st, it = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, 3, 9, mu, P, u, fx["dt"], Q, None, None,
                                       update=False)
        assert st[0] == 0 and it[0] <= 2
        np.testing.assert_allclose(np.trace(P[0][24:36, 24:36]), want[step], rtol=1e-7)
    # the reference then updates on this (indefinite) covariance: LLT fails, status says so (Q8/Q13)
    mu2, P2, st2, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, 3, 9, mu, P, None, 0.0, None, fx["z"][None],
                                     fx["R"], predict=False)
    assert st2[0] & slo.ST_CHOL_FAIL


def test_usckf_clone_order_that_stays_spd(slo):
    """cloning(L); cloning(I) then one predict gives an SPD P (SURVEY 4.3 caveat)."""
    fx = synth.usckf_unit_test_fixture()
    mu = synth.identity_q(synth.STATE_BLOCKS * 3)[None]
    P = np.zeros((1, 36, 36))
    P[0, 24:36, 24:36] = fx["P0_single"]
    P[0, 12:24, 12:24] = fx["P0_single"]
    mu, P = slo.usckf_clone(slo.STATEK_L, 0, 0, mu, P)
    u = np.r_[fx["velo"], fx["angvelo"]][None] * 0.01
    mu, P, st, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, 0, 0, mu, P, u, fx["dt"],
                                  synth.usckf_process_noise(fx["dt"]), None, None, update=False)
    mu, P = slo.usckf_clone(slo.STATEK_I, 0, 0, mu, P)
    mu, P, st, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, 0, 0, mu, P, u, fx["dt"],
                                  synth.usckf_process_noise(fx["dt"]), None, None, update=False)
    assert st[0] == 0
    assert np.linalg.eigvalsh(0.5 * (P[0] + P[0].T)).min() > 0


def test_msckf_remove_outliers_index_quirk(slo):                          # quirk Q6, Msckf.hpp:741-744
    m = 8
    S = np.eye(m)
    innov = np.zeros(m)
    innov[2] = 10.0                      # feature 1 (rows 2,3) is an outlier
    n_out, kept = slo.msckf_remove_outliers(innov, S)
    assert n_out == 1
    # the reference deletes rows {2, 4} (second index not re-based), not {2, 3}
    np.testing.assert_array_equal(kept, [0, 1, 3, 5, 6, 7])
    innov = np.zeros(m)
    innov[6] = 10.0                      # last feature: second deletion falls off the end -> {6,7}
    n_out, kept = slo.msckf_remove_outliers(innov, S)
    assert n_out == 1
    np.testing.assert_array_equal(kept, [0, 1, 2, 3, 4, 5])


def test_check_sigma_points_roundtrip(slo):
    """checkSigmaPoints (Usckf.hpp:769-789): re-estimating from sigma points returns (mu, P)."""
    sc = synth.ukfom_scenario(4, seed=23, p_scale=1e-4)
    zeroQ = np.zeros((9, 9))
    mu, P, st, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, sc["mu"], sc["P"], np.zeros((4, 6)), 0.0, zeroQ,
                                None, None, update=False)
    assert np.max(np.abs(P - sc["P"])) < 1e-6
    for i in range(4):
        assert np.max(np.abs(slo.boxminus([0, 1, 0], mu[i], sc["mu"][i]))) < 1e-12
