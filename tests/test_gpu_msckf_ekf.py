"""GPU parity: Msckf::update, EKF flavour with outlier removal on the information matrix and QR compression
(Msckf.hpp:297-349,756-816) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from parity import STEP_TOL, assert_parity, symmetrize_lower
from slam_localization_b200 import engine, synth

pytestmark = pytest.mark.gpu
BLOCKS = [0, 1, 0, 0] + [0, 1] * 10


def _run(slo, sc, B, gate, k=10):
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], sc["P"])
    f.update_ekf(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"], gate=gate)
    mu_r, P_r, out_r, st_r = slo.msckf_update_ekf(slo.MM_MSCKF_REPROJ, k, sc["mu"], symmetrize_lower(sc["P"]), sc["landmarks"],
                                                  sc["z"], sc["R"], gate=gate)
    return f, mu_r, P_r, out_r, st_r


@pytest.mark.parametrize("B", [1, 5, 160])
def test_ekf_update_no_gate(slo, B):
    sc = synth.msckf_scenario(B, seed=10 + B, k=10, nfeat=50)
    f, mu_r, P_r, out_r, st_r = _run(slo, sc, B, gate=False)
    assert not f.status().any() and not f.outliers().any()
    assert_parity(slo, BLOCKS, f.mu(), symmetrize_lower(f.P()), mu_r, symmetrize_lower(P_r))


# B = 333: more instances than SMs -- from its second instance on a CTA takes the covariance record that was fetched into
# shared memory during the previous instance (also after instances that left early with a status)
@pytest.mark.parametrize("frac,B", [(0.0, 96), (0.04, 96), (0.12, 96), (0.04, 333), (0.12, 333)])
def test_ekf_update_gate_and_outliers(slo, frac, B):
    sc = synth.msckf_scenario(B, seed=77, k=10, nfeat=50, outlier_frac=frac)
    f, mu_r, P_r, out_r, st_r = _run(slo, sc, B, gate=True)
    np.testing.assert_array_equal(f.outliers(), out_r)              # integer outputs: exact
    np.testing.assert_array_equal(f.status(), st_r)
    ok = st_r == 0
    assert ok.any()
    if frac >= 0.1:
        assert (st_r == 16).any()                                   # too few rows left for reduceDimension
    assert_parity(slo, BLOCKS, f.mu(), symmetrize_lower(f.P()), mu_r, symmetrize_lower(P_r), mask=ok)
    # flagged instances are left unchanged
    bad = ~ok
    if bad.any():
        np.testing.assert_array_equal(f.mu()[bad], sc["mu"][bad])
    assert np.linalg.eigvalsh(symmetrize_lower(f.P())[ok]).min() > 0


@pytest.mark.parametrize("k,nfeat", [(9, 50), (4, 21), (7, 43), (10, 37)])
def test_ekf_update_fewer_clones(slo, k, nfeat):
    """Fewer clones / features (k = 9: N = 66): different panel and tile remainders everywhere (m >= N, the QR row rule)."""
    B = 12
    sc = synth.msckf_scenario(B, seed=5, k=k, nfeat=nfeat)
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], sc["P"])
    f.update_ekf(engine.MM_MSCKF_REPROJ, sc["landmarks"], sc["z"], sc["R"], gate=True)
    mu_r, P_r, out_r, st_r = slo.msckf_update_ekf(slo.MM_MSCKF_REPROJ, k, sc["mu"], symmetrize_lower(sc["P"]), sc["landmarks"],
                                                  sc["z"], sc["R"], gate=True)
    np.testing.assert_array_equal(f.outliers(), out_r)
    np.testing.assert_array_equal(f.status(), st_r)
    ok = st_r == 0
    assert_parity(slo, [0, 1, 0, 0] + [0, 1] * k, f.mu(), symmetrize_lower(f.P()), mu_r, symmetrize_lower(P_r), mask=ok)


def test_ekf_update_general_R(slo):
    """A dense (correlated) measurement noise matrix exercises Q^T R Q in full."""
    B = 8
    sc = synth.msckf_scenario(B, seed=9, k=10, nfeat=50)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(100, 100)) * 1e-3
    R = sc["R"] + A @ A.T
    sc = dict(sc, R=0.5 * (R + R.T))
    f, mu_r, P_r, out_r, st_r = _run(slo, sc, B, gate=True)
    np.testing.assert_array_equal(f.outliers(), out_r)
    np.testing.assert_array_equal(f.status(), st_r)
    ok = st_r == 0
    assert_parity(slo, BLOCKS, f.mu(), symmetrize_lower(f.P()), mu_r, symmetrize_lower(P_r), mask=ok)


def test_predict_then_ekf_update_replay(slo):
    """20 predict + EKF-update cycles, both arms free-running from the same start."""
    B = 16
    sc = synth.msckf_scenario(B, seed=21, k=10, nfeat=50)
    f = engine.Msckf(B, nclones=10)
    f.set_state(sc["mu"], sc["P"])
    mu_r, P_r = sc["mu"].copy(), symmetrize_lower(sc["P"])
    u = sc["u"].copy()
    u[:, 0:3] = 0.0
    u[:, 3:7] = [1.0, 0.0, 0.0, 0.0]
    du, dz, dQ, dR, dlm = (engine.DeviceArray(x) for x in (u, sc["z"], sc["Q"], sc["R"], sc["landmarks"]))
    for _ in range(20):
        f.predict(engine.PM_MSCKF_DELTAPOSE, du, 0.0, dQ)
        f.update_ekf(engine.MM_MSCKF_REPROJ, dlm, dz, dR, gate=False)
        mu_r, P_r, _ = slo.msckf_predict(slo.PM_MSCKF_DELTAPOSE, 10, mu_r, P_r, u, 0.0, sc["Q"])
        mu_r, P_r, _, _ = slo.msckf_update_ekf(slo.MM_MSCKF_REPROJ, 10, mu_r, symmetrize_lower(P_r), sc["landmarks"], sc["z"], sc["R"],
                                               gate=False)
    assert not f.status().any()
    assert_parity(slo, BLOCKS, f.mu(), symmetrize_lower(f.P()), mu_r, symmetrize_lower(P_r), tol=1e-7)


def test_ekf_update_diagonal_nonconstant_R(slo):
    """A diagonal but non-constant R takes the QR path with the row-scaling form of Q^T R Q."""
    B = 8
    sc = synth.msckf_scenario(B, seed=11, k=10, nfeat=50)
    rng = np.random.default_rng(1)
    sc = dict(sc, R=np.diag(np.diag(sc["R"]) * rng.uniform(0.5, 2.0, size=100)))
    f, mu_r, P_r, out_r, st_r = _run(slo, sc, B, gate=True)
    np.testing.assert_array_equal(f.outliers(), out_r)
    np.testing.assert_array_equal(f.status(), st_r)
    ok = st_r == 0
    assert ok.any()
    assert_parity(slo, BLOCKS, f.mu(), symmetrize_lower(f.P()), mu_r, symmetrize_lower(P_r), mask=ok)
