"""GPU parity: localization::Msckf (BASELINE config 3: 10 clones, N=72, 50 features, m=100) through the
C ABI vs the CPU oracle."""
import os

import numpy as np
import pytest

import parity
from slam_localization_b200 import engine, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def blocks(k):
    return synth.STATE_BLOCKS + [0, 1] * k


@pytest.mark.parametrize("pm", [engine.PM_MSCKF_DELTAPOSE, engine.PM_USCKF_TEST])
def test_msckf_predict_parity(slo, pm):
    B, k = 77, 10
    sc = synth.msckf_scenario(B, seed=51, k=k)
    u = sc["u"] if pm == engine.PM_MSCKF_DELTAPOSE else sc["u"][:, [0, 1, 2, 10, 11, 12]]
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], sc["P"])
    f.predict(pm, u, 0.01, sc["Q"])
    mu, P, st = slo.msckf_predict(pm, k, sc["mu"], sc["P"], u, 0.01, sc["Q"], nthreads=8)
    assert not st.any() and not f.status().any()
    parity.assert_parity(slo, blocks(k), f.mu(), f.P(), mu, parity.symmetrize_lower(P))
    # cross blocks are left stale (quirk Q5): only the 12x12 corner may change
    Pg = f.P()
    np.testing.assert_array_equal(Pg[:, 12:, :], sc["P"][:, 12:, :])


# (k, nfeat) -> N = 12 + 6k, m = 2 nfeat: every remainder of the 8-wide panels / tiles of the blocked factorisations
# (N mod 8 in {0, 2, 4, 6}, m mod 8 in {0, 2, 4, 6}, m < 8, N < m and N > m)
@pytest.mark.parametrize("k,nfeat", [(10, 50), (3, 6), (4, 7), (1, 1), (2, 9), (5, 10), (7, 11), (9, 33), (6, 49), (8, 3)])
def test_msckf_update_parity_no_outliers(slo, k, nfeat):
    B = 37
    sc = synth.msckf_scenario(B, seed=52, k=k, nfeat=nfeat)
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], sc["P"])
    f.update(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"], gate=False)
    mu, P, out, st, _ = slo.msckf_update(slo.MM_MSCKF_REPROJ, k, sc["mu"], sc["P"], sc["landmarks"], sc["z"], sc["R"],
                                        gate=False, nthreads=8)
    assert not st.any() and not f.status().any()
    parity.assert_parity(slo, blocks(k), f.mu(), f.P(), mu, P)
    Pg = f.P()
    assert np.array_equal(Pg, Pg.transpose(0, 2, 1)) and np.linalg.eigvalsh(Pg).min() > 0


@pytest.mark.parametrize("k,nfeat", [(10, 50), (6, 21), (3, 13)])
def test_msckf_update_with_outliers_matches_reference_quirk(slo, k, nfeat):
    """5% gross outliers: the per-feature 2-dof gate fires and rows are deleted with the reference's index
    quirk (Q6: rows {2i, 2i+2}); the engine must delete the same rows and report the same count."""
    B = 64
    sc = synth.msckf_scenario(B, seed=53, k=k, nfeat=nfeat, outlier_frac=0.05)
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], sc["P"])
    f.update(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"], gate=True)
    mu, P, out, st, _ = slo.msckf_update(slo.MM_MSCKF_REPROJ, k, sc["mu"], sc["P"], sc["landmarks"], sc["z"], sc["R"],
                                        gate=True, nthreads=8)
    assert out.sum() > (20 if nfeat == 50 else 5)
    np.testing.assert_array_equal(f.outliers(), out)
    ok = st == 0
    assert ok.sum() > B // 2
    np.testing.assert_array_equal(f.status()[ok], 0)
    parity.assert_parity(slo, blocks(k), f.mu(), f.P(), mu, P, mask=ok)


def test_msckf_golden_fixture(slo):
    g = np.load(os.path.join(G, "msckf_k10_f50.npz"))
    B = g["mu0"].shape[0]
    f = engine.Msckf(B, nclones=10)
    f.set_state(g["mu0"], g["P0"])
    f.predict(engine.PM_MSCKF_DELTAPOSE, g["u"], 0.0, g["Q"])
    parity.assert_parity(slo, blocks(10), f.mu(), f.P(), g["mu_pred"], parity.symmetrize_lower(g["P_pred"]))
    f.set_state(g["mu0"], g["P0"])
    f.update(engine.MM_MSCKF_REPROJ, g["landmarks"], g["z"], g["R"], gate=True)
    parity.assert_parity(slo, blocks(10), f.mu(), f.P(), g["mu_upd"], g["P_upd"])
    np.testing.assert_array_equal(f.outliers(), g["outliers"])


def test_msckf_indefinite_covariance_is_flagged():
    B, k = 8, 10
    sc = synth.msckf_scenario(B, seed=54, k=k)
    P = sc["P"].copy()
    P[3, 20, 20] = -1.0
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], P)
    f.update(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"])
    st = f.status()
    assert st[3] & engine.ST_CHOL_FAIL and not st[[0, 1, 2, 4, 5, 6, 7]].any()
    np.testing.assert_array_equal(f.mu()[3], sc["mu"][3])


@pytest.mark.parametrize("k,nfeat", [(10, 50), (3, 13), (1, 4)])
def test_msckf_update_several_instances_per_cta(slo, k, nfeat):
    """More instances than SMs: every CTA walks several instances, and from the second one on the covariance factor of
    an instance is computed one iteration early, in lockstep with the previous instance's chol(P_new) (chol_dual).
    Parity against the oracle with the gate on and outliers present, and an indefinite covariance in the second and
    third round of a CTA (flagged when it becomes the 'next' instance; state untouched; neighbours unaffected)."""
    B = 333
    sc = synth.msckf_scenario(B, seed=57, k=k, nfeat=nfeat, outlier_frac=0.04)
    P = sc["P"].copy()
    bad = [160, 161, 310]
    for i in bad:
        P[i, 14, 14] = -1.0
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], P)
    f.update(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"], gate=True)
    mu, Pr, out, st, _ = slo.msckf_update(slo.MM_MSCKF_REPROJ, k, sc["mu"], P, sc["landmarks"], sc["z"], sc["R"],
                                         gate=True, nthreads=8)
    gs = f.status()
    assert all(gs[i] & engine.ST_CHOL_FAIL for i in bad)
    np.testing.assert_array_equal(f.mu()[bad], sc["mu"][bad])
    good = np.ones(B, bool)
    good[bad] = False
    ok = good & (st == 0)
    assert ok.sum() > B // 2
    np.testing.assert_array_equal(gs[ok], 0)
    np.testing.assert_array_equal(f.outliers()[good], out[good])
    parity.assert_parity(slo, blocks(k), f.mu(), f.P(), mu, Pr, mask=ok)


def test_msckf_full_size_oracle_parity(slo):
    """BASELINE configs[2] at its full size: 16,384 instances, 10 clones, 50 features, gate on -- every instance against
    the oracle (<= 1e-9; the fleet is 512 seeded priors replicated with per-instance measurements), outlier counts exact,
    plus: covariances symmetric PSD, status clean, results independent of the batch (bit for bit)."""
    B, k, npri = 16384, 10, 512
    sc = synth.msckf_scenario(npri, seed=55, k=k)
    rep = B // npri
    rng = np.random.default_rng(56)
    z = np.tile(sc["z"], (rep, 1)) + 1e-3 * rng.normal(size=(B, sc["z"].shape[1]))
    z[:npri] = sc["z"]
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], sc["P"], replicate=True)
    f.update(engine.MM_MSCKF_REPROJ, sc["landmarks"], z, sc["R"])
    assert sum(f.status_counts()) == 0
    mu, P = f.mu(), f.P()
    tile = lambda x: np.tile(x, (rep,) + (1,) * (x.ndim - 1))
    mu_r, P_r, out_r, st_r, _ = slo.msckf_update(slo.MM_MSCKF_REPROJ, k, tile(sc["mu"]), tile(sc["P"]), sc["landmarks"], z,
                                                  sc["R"], gate=True, nthreads=16)
    assert not st_r.any()
    np.testing.assert_array_equal(f.outliers(), out_r)
    parity.assert_parity(slo, blocks(k), mu, P, mu_r, P_r)
    g = engine.Msckf(npri, nclones=k)
    g.set_state(sc["mu"], sc["P"])
    g.update(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"])
    np.testing.assert_array_equal(mu[:npri], g.mu())
    np.testing.assert_array_equal(P[:npri], g.P())
    assert np.linalg.eigvalsh(P[::257]).min() > 0


def test_msckf_step_host_matches_device_path():
    B, k = 300, 10
    sc = synth.msckf_scenario(B, seed=91, k=k)
    a, b = engine.Msckf(B, nclones=k), engine.Msckf(B, nclones=k)
    for f in (a, b):
        f.set_state(sc["mu"], sc["P"])
    a.predict(engine.PM_MSCKF_DELTAPOSE, sc["u"], 0.0, sc["Q"])
    a.update(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"])
    out = np.empty((B, 13 + 7 * k))
    b.step_host(engine.PM_MSCKF_DELTAPOSE, engine.MM_MSCKF_REPROJ, sc["u"], 0.0, sc["Q"], np.ascontiguousarray(sc["landmarks"]),
                sc["z"], sc["R"], mu_out=out)
    np.testing.assert_array_equal(out, a.mu())
    np.testing.assert_array_equal(b.P(), a.P())
    np.testing.assert_array_equal(b.outliers(), a.outliers())
