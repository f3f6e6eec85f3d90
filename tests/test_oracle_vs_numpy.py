"""Cross-checks the C++ oracle against an independent numpy/scipy re-implementation (tests/np_ref.py).
The reference asserts no filter output and cannot be built here, so two independent restatements
agreeing far below the 1e-9 parity tolerance is how the oracle is pinned."""
import numpy as np
import pytest

import np_ref
from slam_localization_b200 import synth

TOL = dict(rtol=1e-10, atol=1e-13)


def test_llt_matches_lapack_small_and_blocked(slo):
    rng = np.random.default_rng(3)
    for n in (3, 9, 12, 31, 32, 48, 72, 100):
        P = synth.random_spd(rng, 1, n, scale=0.5, cond=1e4)[0]
        L, info = slo.llt(P)
        assert info == -1
        np.testing.assert_allclose(L, np.linalg.cholesky(P), rtol=1e-11, atol=1e-14)
        assert np.all(np.triu(L, 1) == 0)


def test_llt_reads_lower_triangle_only_and_reports_failure(slo):
    rng = np.random.default_rng(4)
    P = synth.random_spd(rng, 1, 12, 1.0, 10.0)[0]
    Pu = P.copy()
    Pu[np.triu_indices(12, 1)] = 123.0                    # garbage above the diagonal (Q8)
    np.testing.assert_array_equal(slo.llt(Pu)[0], slo.llt(P)[0])
    Pbad = P.copy()
    Pbad[5, 5] = -1.0
    assert slo.llt(Pbad)[1] == 5


def test_inverse_matches_numpy(slo):
    rng = np.random.default_rng(5)
    for n in (2, 3, 6, 12, 100):
        A = synth.random_spd(rng, 1, n, 1.0, 1e3)[0] + 0.1 * rng.normal(size=(n, n))
        np.testing.assert_allclose(slo.inverse(A), np.linalg.inv(A), rtol=1e-9, atol=1e-12)
        if n <= 6:
            np.testing.assert_allclose(slo.inverse(A, fixed=True), np.linalg.inv(A), rtol=1e-9, atol=1e-12)


def test_manifold_ops_match_scipy(slo):
    rng = np.random.default_rng(6)
    blocks = [0, 1, 0, 0, 0, 1]
    for _ in range(20):
        x = synth.random_q(rng, 1, blocks)[0]
        d = rng.normal(size=18) * rng.choice([1e-6, 1e-2, 0.5, 1.5])
        y = slo.boxplus(blocks, x, d)
        yn = np_ref.boxplus(blocks, x, d)
        # quaternions agree up to sign
        for o in (3, 16):
            if np.dot(y[o:o + 4], yn[o:o + 4]) < 0:
                yn[o:o + 4] *= -1
        np.testing.assert_allclose(y, yn, rtol=1e-12, atol=1e-14)
        x2 = synth.random_q(rng, 1, blocks, max_angle=1.0)[0]
        x3 = slo.boxplus(blocks, x2, rng.normal(size=18) * 0.5)
        np.testing.assert_allclose(slo.boxminus(blocks, x3, x2), np_ref.boxminus(blocks, x3, x2), rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("layout,pm", [(9, 1), (9, 2), (6, 3)])
def test_ukf_predict_update_vs_numpy(slo, layout, pm):
    sc = synth.ukfom_scenario(6, seed=11, layout=layout)
    blocks = synth.LAYOUT_BLOCKS[layout]
    mu1, P1, st, _ = slo.ukf_step(layout, pm, slo.MM_GPS_POS, sc["mu"], sc["P"], sc["u"], sc["dt"], sc["Q"],
                                  None, None, update=False)
    mu2, P2, st2, _ = slo.ukf_step(layout, pm, slo.MM_GPS_POS, mu1, P1, None, 0.0, None, sc["z"], sc["R"],
                                   predict=False)
    assert not st.any() and not st2.any()
    for i in range(6):
        if pm == 3:
            g = lambda s: np_ref.pm_pose6_odom(s, sc["u"][i], sc["dt"])
        else:
            g = lambda s: np_ref.pm_ukfom_imu(s, sc["u"][i], sc["dt"], refbug=(pm == 2))
        m_ref, P_ref = np_ref.ukf_predict(blocks, sc["mu"][i], sc["P"][i], g, sc["Q"])
        assert np.max(np.abs(np_ref.boxminus(blocks, mu1[i], m_ref))) < 1e-12
        np.testing.assert_allclose(P1[i], P_ref, **TOL)
        m_ref2, P_ref2 = np_ref.ukf_update(blocks, m_ref, P_ref, sc["z"][i], lambda s: s[0:3], sc["R"])
        assert np.max(np.abs(np_ref.boxminus(blocks, mu2[i], m_ref2))) < 1e-11
        np.testing.assert_allclose(P2[i], P_ref2, rtol=1e-9, atol=1e-13)


def test_usckf_predict_update_vs_numpy(slo):
    sc = synth.usckf_scenario(4, seed=12)
    nk, nl = sc["nk"], sc["nl"]
    assert np.linalg.eigvalsh(sc["P"]).min() > 0
    mu1, P1, st, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, nk, nl, sc["mu"], sc["P"], sc["u"],
                                    sc["dt"], sc["Q"], None, None, update=False)
    mu2, P2, st2, _ = slo.usckf_step(slo.PM_USCKF_TEST, slo.MM_USCKF_VO, nk, nl, mu1, P1, None, 0.0, None,
                                     sc["z"], sc["R"], predict=False)
    assert not st.any() and not st2.any()
    blocks = np_ref.AUG_BLOCKS
    for i in range(4):
        f = lambda s: np_ref.pm_usckf_test(s, sc["u"][i], sc["dt"])
        m_ref, P_ref = np_ref.usckf_predict(sc["mu"][i], sc["P"][i], nk, nl, f, sc["Q"])
        assert np.max(np.abs(np_ref.boxminus(blocks, mu1[i], m_ref, nk + nl))) < 1e-12
        np.testing.assert_allclose(P1[i], P_ref, rtol=1e-9, atol=1e-14)
        m_ref2, P_ref2 = np_ref.usckf_update(m_ref, P_ref, nk, nl, sc["z"][i], sc["R"])
        assert np.max(np.abs(np_ref.boxminus(blocks, mu2[i], m_ref2, nk + nl))) < 1e-11
        np.testing.assert_allclose(P2[i], P_ref2, rtol=1e-9, atol=1e-14)


def test_msckf_predict_update_vs_numpy(slo):
    k, nfeat = 3, 6
    sc = synth.msckf_scenario(2, seed=13, k=k, nfeat=nfeat)
    blocks = np_ref.multi_blocks(k)
    mu1, P1, st = slo.msckf_predict(slo.PM_MSCKF_DELTAPOSE, k, sc["mu"], sc["P"], sc["u"], 0.0, sc["Q"])
    mu2, P2, out, st2, _ = slo.msckf_update(slo.MM_MSCKF_REPROJ, k, mu1, P1, sc["landmarks"], sc["z"], sc["R"],
                                           gate=False)
    assert not st.any() and not st2.any() and not out.any()
    for i in range(2):
        f = lambda s: np_ref.pm_msckf_deltapose(s, sc["u"][i])
        m_s, P_s = np_ref.ukf_predict(np_ref.STATE_BLOCKS, sc["mu"][i][:13], sc["P"][i][:12, :12], f, sc["Q"])
        m_ref = sc["mu"][i].copy()
        m_ref[:13] = m_s
        P_ref = sc["P"][i].copy()
        P_ref[:12, :12] = P_s                                # cross blocks stay stale (Q5)
        assert np.max(np.abs(np_ref.boxminus(blocks, mu1[i], m_ref))) < 1e-12
        np.testing.assert_allclose(P1[i], P_ref, **TOL)
        h = lambda s: np_ref.mm_msckf_reproj(s, k, sc["landmarks"])
        m_ref2, P_ref2 = np_ref.ukf_update(blocks, m_ref, P_ref, sc["z"][i], h, sc["R"])
        assert np.max(np.abs(np_ref.boxminus(blocks, mu2[i], m_ref2))) < 1e-10
        np.testing.assert_allclose(P2[i], P_ref2, rtol=1e-8, atol=1e-13)


def test_fusion_vs_numpy(slo):
    for d in (3, 6):
        # moderately conditioned: the two implementations agree to ~1e-10
        sc = synth.fusion_scenario(16, d=d, log_spread=0.5)
        xo, Co = slo.datamodel(0, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
        for i in range(16):
            x_ref, C_ref = np_ref.fusion(sc["x1"][i], sc["C1"][i], sc["x2"][i], sc["C2"][i])
            np.testing.assert_allclose(xo[i], x_ref, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(Co[i], C_ref, rtol=1e-9, atol=1e-14)
        # SURVEY config-5 conditioning (cond up to ~1e7): explicit inverses lose ~cond*eps, so two
        # correct implementations only agree to that level -- bounded here norm-wise
        sc = synth.fusion_scenario(64, d=d)
        xo, Co = slo.datamodel(0, sc["x1"], sc["C1"], sc["x2"], sc["C2"])
        for i in range(64):
            x_ref, C_ref = np_ref.fusion(sc["x1"][i], sc["C1"][i], sc["x2"][i], sc["C2"][i])
            kappa = max(np.linalg.cond(sc["C1"][i]), np.linalg.cond(sc["C2"][i]))
            assert np.linalg.norm(xo[i] - x_ref) <= 1e-13 * kappa * max(1.0, np.linalg.norm(x_ref))
            assert np.linalg.norm(Co[i] - C_ref) <= 1e-13 * kappa * np.linalg.norm(C_ref)
