"""GPU parity for the SURVEY 8(f) "next" rows through the C ABI vs the CPU oracle:
f2 error-state EKF (ekfPredict / ekfUpdate Joseph form / ekfSingleUpdate / cloning, UsckfError.hpp),
f3 DataModel::safeFusion, f4 DeadReckon::updatePose + TransformWithUncertainty::operator*.
Tolerance: relative 1e-9 (north_star) for the floating-point paths; integer outputs (gate decisions) exact."""
import numpy as np
import pytest

from parity import STEP_TOL, cov_error
from slam_localization_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1000])
def test_ekf_predict(slo, n):
    sc = synth.ekf_scenario(n, seed=n)
    f = engine.ErrorStateEkf(sc["mu"], sc["err"], sc["P"])
    f.ekf_predict(sc["F"], sc["Q"])
    err_r, P_r = slo.ekf_predict(sc["err"], sc["P"], sc["F"], sc["Q"])
    assert cov_error(f.P.numpy(), P_r) <= STEP_TOL
    assert _rel(f.err.numpy(), err_r) <= STEP_TOL
    np.testing.assert_array_equal(f.mu.numpy(), sc["mu"])


@pytest.mark.parametrize("n", [1, 8, 33, 2000])
@pytest.mark.parametrize("gate", [False, True])
def test_ekf_update_joseph(slo, n, gate):
    sc = synth.ekf_scenario(n, seed=100 + n, outlier_frac=0.2 if gate else 0.0)
    f = engine.ErrorStateEkf(sc["mu"], sc["err"], sc["P"])
    ret = f.ekf_update(sc["z"], sc["H"], sc["R"], gate=gate).numpy()
    P_r, ret_r, acc_r = slo.ekf_update(sc["mu"], sc["P"], sc["z"], sc["H"], sc["R"], gate=1 if gate else 0)
    acc = f.accepted.cpu().numpy()
    np.testing.assert_array_equal(acc, acc_r)
    if gate and n >= 33:
        assert 0 < acc.sum() < n
    P = f.P.numpy()
    assert cov_error(P, P_r) <= STEP_TOL
    assert _rel(ret, ret_r) <= STEP_TOL if ret_r.any() else not ret.any()
    np.testing.assert_array_equal(P, P.transpose(0, 2, 1))            # symmetry guaranteed (:359)
    np.testing.assert_array_equal(P[acc == 0], sc["P"][acc == 0])      # rejected: untouched
    assert np.linalg.eigvalsh(P).min() > 0                             # PSD preserved
    np.testing.assert_array_equal(f.mu.numpy(), sc["mu"])             # only Pk_error changes (:354 is a local)


@pytest.mark.parametrize("n", [1, 8, 1001])
def test_ekf_single_update_and_clone(slo, n):
    sc = synth.ekf_scenario(n, seed=7 + n)
    f = engine.ErrorStateEkf(sc["mu"], sc["err"], sc["P"])
    f.ekf_single_update(sc["zs"], sc["Hs"], sc["R"], gate=False)
    mu_r, P_r, acc_r = slo.ekf_single_update(sc["mu"], sc["err"], sc["P"], sc["zs"], sc["Hs"], sc["R"], gate=0)
    np.testing.assert_array_equal(f.accepted.cpu().numpy(), acc_r)
    assert cov_error(f.P.numpy(), P_r) <= STEP_TOL
    assert _rel(f.mu.numpy(), mu_r) <= STEP_TOL
    f.cloning()
    mu_c, err_c, P_c = slo.ekf_clone(mu_r, sc["err"], P_r)
    assert cov_error(f.P.numpy(), P_c) <= STEP_TOL
    assert _rel(f.mu.numpy(), mu_c) <= STEP_TOL
    np.testing.assert_array_equal(f.err.numpy(), err_c)
    # gated single update: rejected instances still receive the correction from mu_error (:553-568)
    g = engine.ErrorStateEkf(sc["mu"], sc["err"], sc["P"])
    zs = sc["zs"] + 3.0
    g.ekf_single_update(zs, sc["Hs"], sc["R"], gate=True)
    mu_r, P_r, acc_r = slo.ekf_single_update(sc["mu"], sc["err"], sc["P"], zs, sc["Hs"], sc["R"], gate=1)
    assert not acc_r.any()
    np.testing.assert_array_equal(g.accepted.cpu().numpy(), acc_r)
    assert _rel(g.mu.numpy(), mu_r) <= STEP_TOL
    np.testing.assert_array_equal(g.P.numpy(), sc["P"])


def test_ekf_predict_update_replay(slo):
    """200 free-running predict/update cycles from the same start: drift between the two arms stays tiny."""
    n = 64
    sc = synth.ekf_scenario(n, seed=42)
    f = engine.ErrorStateEkf(sc["mu"], sc["err"], sc["P"])
    err_r, P_r = sc["err"].copy(), sc["P"].copy()
    F, Q, z, H, R = (engine.DeviceArray(sc[k]) for k in ("F", "Q", "z", "H", "R"))
    for _ in range(200):
        f.ekf_predict(F, Q)
        f.ekf_update(z, H, R, gate=False)
        err_r, P_r = slo.ekf_predict(err_r, P_r, sc["F"], sc["Q"])
        P_r, _, _ = slo.ekf_update(sc["mu"], P_r, sc["z"], sc["H"], sc["R"], gate=0)
    P = f.P.numpy()
    assert cov_error(P, P_r) <= 1e-7
    assert np.linalg.eigvalsh(P).min() > 0


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000])
def test_safe_fusion(slo, n):
    sc = synth.safe_fusion_scenario(n, log_spread=1.5)
    xo, Co = engine.DataModel.safe_fuse(sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    xr, Cr = slo.safe_fusion(sc["x1"], sc["C1"], sc["x2"], sc["C2"])
    xo, Co = xo.numpy(), Co.numpy()
    # same operation order, no FMA contraction, IEEE sqrt/div on both sides
    np.testing.assert_array_equal(xo, xr)
    np.testing.assert_array_equal(Co, Cr)


def test_safe_fusion_reference_inputs(slo):
    fx = synth.safe_fusion_fixture()                       # test/DataModelUnitTest.cpp:66-74
    xo, Co = engine.DataModel.safe_fuse(fx["x1"], fx["C1"], fx["x2"], fx["C2"])
    xr, Cr = slo.safe_fusion(fx["x1"], fx["C1"], fx["x2"], fx["C2"])
    np.testing.assert_array_equal(xo.numpy(), xr)
    np.testing.assert_array_equal(Co.numpy(), Cr)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 3000])
def test_transform_compose(slo, n):
    rng = np.random.default_rng(n)
    p2 = np.concatenate([rng.normal(size=(n, 3)), synth.random_unit_quat(rng, n)], axis=1)
    p1 = np.concatenate([rng.normal(size=(n, 3)), synth.random_unit_quat(rng, n)], axis=1)
    p1[::3, 3:] *= -1.0
    c2 = synth.random_spd(rng, n, 6, scale=1e-2, cond=1e3)
    c1 = synth.random_spd(rng, n, 6, scale=1e-2, cond=1e3)
    po, co = engine.DeadReckon.compose(p2, c2, p1, c1)
    pr, cr = slo.transform_compose(p2, c2, p1, c1)
    assert _rel(po.numpy(), pr) <= STEP_TOL
    assert cov_error(co.numpy(), cr) <= STEP_TOL


@pytest.mark.parametrize("n", [1, 64, 4097])
def test_dead_reckon_update_pose(slo, n):
    sc = synth.deadreckon_scenario(n, seed=n)
    out = engine.DeadReckon.update_pose(sc["dt"], sc["vel0"], sc["vel1"], sc["velcov"], sc["prev_pose"], sc["prev_cov"])
    ref = slo.dr_update_pose(sc["dt"], sc["vel0"], sc["vel1"], sc["velcov"], sc["prev_pose"], sc["prev_cov"])
    names = ("post_pose", "post_cov", "delta_pose", "delta_cov")
    for nm, a, b in zip(names, out, ref):
        a = a.numpy()
        if a.ndim == 3:
            assert cov_error(a, b) <= STEP_TOL, nm
        else:
            assert _rel(a, b) <= STEP_TOL, nm


def test_dead_reckon_chain(slo):
    """300 chained updatePose steps (the odometry front-end feeding Msckf's process model).  The reference never
    re-normalises the quaternion it recovers from R2 R1, so its norm drifts -- identically in both arms."""
    n = 32
    sc = synth.deadreckon_scenario(n, seed=3)
    pose, cov = sc["prev_pose"].copy(), sc["prev_cov"].copy()
    dpose, dcov = engine.DeviceArray(pose), engine.DeviceArray(cov)
    vel0, vel1, velcov = (engine.DeviceArray(sc[k]) for k in ("vel0", "vel1", "velcov"))
    for _ in range(300):
        dpose, dcov, _, _ = engine.DeadReckon.update_pose(sc["dt"], vel0, vel1, velcov, dpose, dcov)
        pose, cov, _, _ = slo.dr_update_pose(sc["dt"], sc["vel0"], sc["vel1"], sc["velcov"], pose, cov)
    a = dpose.numpy()
    # compare rotations, not quaternion signs
    assert _rel(a[:, :3], pose[:, :3]) <= 1e-7
    sgn = np.sign(np.sum(a[:, 3:] * pose[:, 3:], axis=1, keepdims=True))
    assert _rel(sgn * a[:, 3:], pose[:, 3:]) <= 1e-7
    assert cov_error(dcov.numpy(), cov) <= 1e-6


def test_next_rows_against_committed_golden():
    """The CUDA path against the committed fixtures of tests/golden/ (no oracle in the loop)."""
    import os
    G = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(G, "ekf_n45.npz"))
    f = engine.ErrorStateEkf(g["mu0"], g["err0"], g["P0"])
    f.ekf_predict(g["F"], g["Q"])
    assert cov_error(f.P.numpy(), g["P1"]) <= STEP_TOL
    ret = f.ekf_update(g["z"], g["H"], g["R"], gate=True).numpy()
    np.testing.assert_array_equal(f.accepted.cpu().numpy(), g["acc"])
    assert cov_error(f.P.numpy(), g["P2"]) <= STEP_TOL
    assert _rel(ret, g["ret"]) <= STEP_TOL
    f.ekf_single_update(g["zs"], g["Hs"], g["R"], gate=True)
    np.testing.assert_array_equal(f.accepted.cpu().numpy(), g["acc3"])
    assert cov_error(f.P.numpy(), g["P3"]) <= STEP_TOL
    assert _rel(f.mu.numpy(), g["mu3"]) <= STEP_TOL
    g = np.load(os.path.join(G, "safe_fusion_d3.npz"))
    xo, Co = engine.DataModel.safe_fuse(g["x1"], g["C1"], g["x2"], g["C2"])
    np.testing.assert_array_equal(xo.numpy(), g["xo"])
    np.testing.assert_array_equal(Co.numpy(), g["Co"])
    g = np.load(os.path.join(G, "deadreckon.npz"))
    out = engine.DeadReckon.update_pose(float(g["dt"]), g["vel0"], g["vel1"], g["velcov"], g["prev_pose"], g["prev_cov"])
    for a, k in zip(out, ("post", "pcov", "dpose", "dcov")):
        assert _rel(a.numpy(), g[k]) <= STEP_TOL, k


def test_ekf_full_size_properties(slo):
    """SURVEY 8f row f2 at its bench size (262,144 instances): Joseph update keeps every covariance exactly symmetric
    and PSD; a strided sample must match the oracle; instances are independent of the batch they sit in."""
    n, npri = 262144, 1024
    sc = synth.ekf_scenario(npri, seed=91)
    rep = n // npri
    f = engine.ErrorStateEkf(np.tile(sc["mu"], (rep, 1)), np.tile(sc["err"], (rep, 1)), np.tile(sc["P"], (rep, 1, 1)))
    F, z = np.tile(sc["F"], (rep, 1, 1)), np.tile(sc["z"], (rep, 1))
    f.ekf_predict(F, sc["Q"])
    f.ekf_update(z, sc["H"], sc["R"], gate=True)
    assert int(f.accepted.sum().item()) == rep * int(slo.ekf_update(sc["mu"], slo.ekf_predict(sc["err"], sc["P"], sc["F"], sc["Q"])[1],
                                                                    sc["z"], sc["H"], sc["R"], gate=1)[2].sum())
    P = f.P.t[::4099].cpu().numpy()
    np.testing.assert_array_equal(P, P.transpose(0, 2, 1))
    assert np.linalg.eigvalsh(P).min() > 0
    # replicas of the same prior give bitwise the same posterior wherever they sit in the batch
    Pall = f.P.t
    assert bool((Pall[:npri] == Pall[-npri:]).all().item())
    err_r, P_r = slo.ekf_predict(sc["err"], sc["P"], sc["F"], sc["Q"])
    P_r, _, _ = slo.ekf_update(sc["mu"], P_r, sc["z"], sc["H"], sc["R"], gate=1)
    assert cov_error(Pall[:npri].cpu().numpy(), P_r) <= STEP_TOL


def test_msckf_ekf_full_size(slo):
    """SURVEY 8f row f1 at the config-3 size (16,384 instances, 10 clones, 50 features)."""
    from parity import assert_parity, symmetrize_lower
    B, npri, k = 16384, 256, 10
    sc = synth.msckf_scenario(npri, seed=92, k=k, nfeat=50, outlier_frac=0.03)
    f = engine.Msckf(B, nclones=k)
    f.set_state(sc["mu"], sc["P"], replicate=True)
    z = np.tile(sc["z"], (B // npri, 1))
    f.update_ekf(engine.MM_MSCKF_REPROJ, sc["landmarks"], z, sc["R"], gate=True)
    mu_r, P_r, out_r, st_r = slo.msckf_update_ekf(slo.MM_MSCKF_REPROJ, k, sc["mu"], symmetrize_lower(sc["P"]), sc["landmarks"], sc["z"],
                                                  sc["R"], gate=True, nthreads=8)
    out, st = f.outliers(), f.status()
    np.testing.assert_array_equal(out.reshape(-1, npri), np.tile(out_r, (B // npri, 1)))   # every replica agrees with the oracle
    np.testing.assert_array_equal(st.reshape(-1, npri), np.tile(st_r, (B // npri, 1)))
    ok = st_r == 0
    mu, P = f.mu(first=npri), f.P(first=npri)
    assert_parity(slo, [0, 1, 0, 0] + [0, 1] * k, mu, symmetrize_lower(P), mu_r, symmetrize_lower(P_r), mask=ok)
    mu_all = f.mu()
    np.testing.assert_array_equal(mu_all[:npri], mu_all[-npri:])
