"""GPU parity: batched ukfom::ukf (config 2 of BASELINE.json) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

import parity
from slam_localization_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _run_gpu(sc, layout, pm, fused, B):
    f = engine.Ukf(B, layout=layout)
    f.set_state(sc["mu"], sc["P"])
    if fused:
        f.step(pm, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"])
        return f, None
    f.predict(pm, sc["u"], sc["dt"], sc["Q"])
    mid = (f.mu(), f.P())
    f.update(engine.MM_GPS_POS, sc["z"], sc["R"])
    return f, mid


@pytest.mark.parametrize("layout,pm", [(9, engine.PM_UKFOM_IMU), (9, engine.PM_UKFOM_IMU_REFBUG), (6, engine.PM_POSE6_ODOM)])
@pytest.mark.parametrize("fused", [False, True])
def test_ukf_step_parity(slo, layout, pm, fused):
    B = 1000                                   # ragged: not a multiple of the CTA size
    sc = synth.ukfom_scenario(B, seed=31, layout=layout)
    blocks = synth.LAYOUT_BLOCKS[layout]
    f, mid = _run_gpu(sc, layout, pm, fused, B)
    mu1, P1, st1, _ = slo.ukf_step(layout, pm, slo.MM_GPS_POS, sc["mu"], sc["P"], sc["u"], sc["dt"], sc["Q"], None,
                                   None, update=False)
    if mid is not None:
        parity.assert_parity(slo, blocks, mid[0], mid[1], mu1, P1)
    mu2, P2, st2, _ = slo.ukf_step(layout, pm, slo.MM_GPS_POS, mu1, P1, None, 0.0, None, sc["z"], sc["R"], predict=False)
    em, ec = parity.assert_parity(slo, blocks, f.mu(), f.P(), mu2, P2)
    assert not f.status().any() and not st2.any()
    P = f.P()
    assert np.array_equal(P, P.transpose(0, 2, 1))
    assert np.linalg.eigvalsh(P).min() > 0


def test_ukf_reference_fixture(slo):
    """The reference's own UKFOM test inputs (test/UKFoMUnitTest.cpp:95-114), batch of one."""
    fx = synth.ukfom_fixture()
    f = engine.Ukf(1)
    f.set_state(fx["mu"], fx["P"])
    f.predict(engine.PM_UKFOM_IMU_REFBUG, fx["u"], fx["dt"], fx["Q"])
    f.update(engine.MM_GPS_POS, fx["z"], fx["R"])
    mu, P, st, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU_REFBUG, slo.MM_GPS_POS, fx["mu"], fx["P"], fx["u"], fx["dt"], fx["Q"],
                                fx["z"], fx["R"])
    parity.assert_parity(slo, [0, 1, 0], f.mu(), f.P(), mu, P)


def test_ukf_golden_fixture(slo):
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ukf_mtk9.npz"))
    B = g["mu0"].shape[0]
    f = engine.Ukf(B)
    f.set_state(g["mu0"], g["P0"])
    f.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, g["u"], float(g["dt"]), g["Q"], g["z"], g["R"])
    parity.assert_parity(slo, [0, 1, 0], f.mu(), f.P(), g["mu1"], g["P1"])


def test_ukf_gate_and_failure_flags(slo):
    B = 64
    sc = synth.ukfom_scenario(B, seed=32)
    sc["z"][:8] += 50.0                         # far outside the 3-dof 5% gate
    P = sc["P"].copy()
    P[8, 4, 4] = -1.0                           # indefinite: LLT must flag, instance left untouched
    f = engine.Ukf(B)
    f.set_state(sc["mu"], P)
    f.update(engine.MM_GPS_POS, sc["z"], sc["R"], gate_dof=3)
    st = f.status()
    assert np.all(st[:8] & engine.ST_GATE_REJECT)
    assert st[8] & engine.ST_CHOL_FAIL
    mu, Pg = f.mu(), f.P()
    np.testing.assert_array_equal(mu[:9], sc["mu"][:9])                      # rejected: state untouched
    np.testing.assert_array_equal(np.tril(Pg[:9]), np.tril(P[:9]))
    mu_r, P_r, st_r, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, sc["mu"], P, None, 0.0, None, sc["z"], sc["R"],
                                      gate_dof=3, predict=False)
    ok = st_r == 0
    assert ok.sum() >= B - 9 - 4
    np.testing.assert_array_equal(st[ok], 0)
    parity.assert_parity(slo, [0, 1, 0], mu, Pg, mu_r, P_r, mask=ok)
    assert f.status_counts()[0] == 1 and f.status_counts()[2] >= 8
    f.clear_status()
    assert not f.status().any()


def test_ukf_free_running_10000_steps(slo):
    """north_star: relative error <= 1e-6 on means and covariances after 10k free-running steps (GPU and oracle each
    keep filtering their own state from the same IMU / GPS stream), covariance symmetric PSD at the end, and the
    filter actually tracks the simulated truth.  The stream is synth.UkfomTruthRun: an observable, bounded flight (see
    its docstring for why a replay drawn around the estimator's own mean cannot be used for a 10k-step comparison)."""
    B, steps = 16, 10000
    run = synth.UkfomTruthRun(B, seed=33)
    mu, P = run.initial(p_scale=1e-4)
    Q, R = synth.ukfom_process_noise(run.dt), (run.r_sigma ** 2) * np.eye(3)
    f = engine.Ukf(B)
    f.set_state(mu, P)
    for k in range(steps):
        u, z = run.step()
        f.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, u, run.dt, Q, z, R)
        mu, P, st, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, mu, P, u, run.dt, Q, z, R, nthreads=4)
        assert not st.any()
    Pg = f.P()
    parity.assert_parity(slo, [0, 1, 0], f.mu(), Pg, mu, P, tol=parity.LONG_TOL)
    assert not f.status().any()
    assert np.array_equal(Pg, Pg.transpose(0, 2, 1)) and np.linalg.eigvalsh(Pg).min() > 0
    assert np.abs(f.mu()[:, :3] - run.p).max() < 0.05            # 1 cm GPS noise: centimetre-level tracking


def test_ukf_full_size_oracle_parity(slo):
    """BASELINE configs[1] at its full size: all 65,536 instances of one fused predict+update step against the oracle
    (<= 1e-9), plus the size-independent properties: identity process model returns P + Q (checkSigmaPoints,
    Usckf.hpp:769-789), symmetric PSD covariances, instance i of the big batch equals instance i run alone bit for bit."""
    B = 65536
    sc = synth.ukfom_scenario(B, seed=34, p_scale=1e-4)
    f = engine.Ukf(B)
    f.set_state(sc["mu"], sc["P"])
    f.predict(engine.PM_UKFOM_IMU, np.zeros((B, 6)), 0.0, sc["Q"])
    P = f.P()
    assert np.max(np.abs(P - (sc["P"] + sc["Q"]))) / np.max(np.abs(sc["P"])) < 1e-9
    f.set_state(sc["mu"], sc["P"])
    f.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"])
    mu_big, P_big = f.mu(), f.P()
    assert not f.status().any()
    mu_r, P_r, st_r, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, sc["mu"], sc["P"], sc["u"], sc["dt"], sc["Q"],
                                      sc["z"], sc["R"], nthreads=16)
    assert not st_r.any()
    parity.assert_parity(slo, [0, 1, 0], mu_big, P_big, mu_r, P_r)
    assert np.linalg.eigvalsh(P_big[::97]).min() > 0
    idx = np.array([0, 1, 31, 32, 127, 128, 4095, 65535])
    g = engine.Ukf(len(idx))
    g.set_state(sc["mu"][idx], sc["P"][idx])
    g.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"][idx], sc["dt"], sc["Q"], sc["z"][idx], sc["R"])
    np.testing.assert_array_equal(g.mu(), mu_big[idx])
    np.testing.assert_array_equal(g.P(), P_big[idx])


def test_ukf_step_host_matches_device_path():
    B = 512
    sc = synth.ukfom_scenario(B, seed=35)
    a, b = engine.Ukf(B), engine.Ukf(B)
    a.set_state(sc["mu"], sc["P"])
    b.set_state(sc["mu"], sc["P"])
    a.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"])
    out = np.empty((B, 10))
    b.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"], mu_out=out)
    np.testing.assert_array_equal(out, a.mu())
    np.testing.assert_array_equal(b.P(), a.P())


def test_ensemble_stats_matches_host(slo):
    B = 3000
    sc = synth.ukfom_scenario(B, seed=36)
    f = engine.Ukf(B)
    f.set_state(sc["mu"], sc["P"])
    out = f.ensemble_stats().numpy()
    X = np.array([slo.get_vectorized([0, 1, 0], m) for m in sc["mu"]])
    assert out[0] == B
    np.testing.assert_allclose(out[1:10], X.sum(0), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(out[10:].reshape(9, 9), X.T @ X, rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("B", [40000, 65536 + 17])
def test_ukf_step_host_chunked_pipeline_is_bitwise_the_device_path(B):
    """slb_ukf_step_host cuts batches >= 32768 into chunks on internal streams (H2D / kernel / D2H overlap);
    ragged sizes must give exactly what the single-launch device path gives."""
    sc = synth.ukfom_scenario(B, seed=37)
    a, b = engine.Ukf(B), engine.Ukf(B)
    a.set_state(sc["mu"], sc["P"])
    b.set_state(sc["mu"], sc["P"])
    a.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"])
    out = np.empty((B, 10))
    b.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], sc["R"], mu_out=out)
    np.testing.assert_array_equal(out, a.mu())
    np.testing.assert_array_equal(b.P(), a.P())
    np.testing.assert_array_equal(b.status(), a.status())


def test_ukf_step_host_graph_replay_with_pinned_buffers():
    """With page-locked host buffers slb_ukf_step_host captures its H2D / kernel / D2H pipeline into a CUDA graph and
    replays it while the same buffers are passed: three steps (capture + two replays, new inputs written into the same
    pinned buffers) must equal three device-path steps bit for bit; changing a parameter re-captures."""
    import torch
    B = 40000
    sc = synth.ukfom_scenario(B, seed=38)
    a, b = engine.Ukf(B), engine.Ukf(B)
    a.set_state(sc["mu"], sc["P"])
    b.set_state(sc["mu"], sc["P"])
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    hu, hz, hQ, hR = pin(sc["u"]), pin(sc["z"]), pin(sc["Q"]), pin(sc["R"])
    hout = torch.empty((B, 10), dtype=torch.float64).pin_memory()
    n0 = engine.launch_count()
    for k in range(3):
        u, z = sc["u"] * (1.0 + 0.1 * k), sc["z"] + 0.01 * k
        hu.copy_(torch.from_numpy(u))
        hz.copy_(torch.from_numpy(z))
        a.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, u, sc["dt"], sc["Q"], z, sc["R"])
        b.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, hu, sc["dt"], hQ, hz, hR, mu_out=hout)
        np.testing.assert_array_equal(hout.numpy(), a.mu())
    np.testing.assert_array_equal(b.P(), a.P())
    assert engine.launch_count() - n0 >= 3 * 3        # replays count the kernels they launch
    # a different gate re-captures and still matches
    a.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"] + 5.0, sc["R"], gate_dof=3)
    hu.copy_(torch.from_numpy(sc["u"]))
    hz.copy_(torch.from_numpy(sc["z"] + 5.0))
    b.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, hu, sc["dt"], hQ, hz, hR, gate_dof=3, mu_out=hout)
    np.testing.assert_array_equal(hout.numpy(), a.mu())
    np.testing.assert_array_equal(b.status(), a.status())
    assert (a.status() & engine.ST_GATE_REJECT).any()


def test_ukf_step_host_async_pipeline_equals_synchronous_steps():
    import torch
    B = 20000
    sc = synth.ukfom_scenario(B, seed=91)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    rng = np.random.default_rng(92)
    us = [pin(sc["u"] + 0.01 * rng.normal(size=sc["u"].shape)) for _ in range(3)]
    zs = [pin(sc["z"] + 0.01 * rng.normal(size=sc["z"].shape)) for _ in range(3)]
    hQ, hR = pin(sc["Q"]), pin(sc["R"])
    a, b = engine.Ukf(B), engine.Ukf(B)
    for f in (a, b):
        f.set_state(sc["mu"], sc["P"])
    oa = [torch.empty((B, 10), dtype=torch.float64).pin_memory() for _ in range(3)]
    ob = [torch.empty((B, 10), dtype=torch.float64).pin_memory() for _ in range(3)]
    for k in range(3):
        a.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, us[k], sc["dt"], hQ, zs[k], hR, mu_out=oa[k])
    for k in range(3):
        b.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, us[k], sc["dt"], hQ, zs[k], hR, mu_out=ob[k], wait=False)
    b.wait()
    for k in range(3):
        np.testing.assert_array_equal(oa[k].numpy(), ob[k].numpy())
    np.testing.assert_array_equal(a.P(), b.P())


def test_ukf_fused_step_keeps_the_predicted_state_when_the_update_fails(slo):
    """step() must equal predict() followed by update() also when the update cannot go through.  A negative
    definite R (invalid input) makes P - K S K^T indefinite: the post-update Cholesky fails, the instance is flagged and
    keeps its PREDICTED state -- in the fused launch exactly as in the two-call sequence and in the oracle."""
    B = 200
    sc = synth.ukfom_scenario(B, seed=36, p_scale=1e-4)
    Rbad = -2e-5 * np.eye(3)
    a, b = engine.Ukf(B), engine.Ukf(B)
    for f in (a, b):
        f.set_state(sc["mu"], sc["P"])
    a.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], sc["Q"], sc["z"], Rbad)
    b.predict(engine.PM_UKFOM_IMU, sc["u"], sc["dt"], sc["Q"])
    mu_p, P_p = b.mu(), b.P()
    b.update(engine.MM_GPS_POS, sc["z"], Rbad)
    failed = (b.status() & engine.ST_CHOL_FAIL) != 0
    assert failed.sum() > B // 2
    np.testing.assert_array_equal(a.status(), b.status())
    np.testing.assert_array_equal(a.mu(), b.mu())
    np.testing.assert_array_equal(a.P(), b.P())
    np.testing.assert_array_equal(b.mu()[failed], mu_p[failed])          # the failed update left the predicted state alone
    np.testing.assert_array_equal(b.P()[failed], P_p[failed])
    mu_r, P_r, st_r, _ = slo.ukf_step(9, slo.PM_UKFOM_IMU, slo.MM_GPS_POS, sc["mu"], sc["P"], sc["u"], sc["dt"], sc["Q"],
                                      None, None, update=False, nthreads=4)
    parity.assert_parity(slo, [0, 1, 0], a.mu()[failed], a.P()[failed], mu_r[failed], P_r[failed])


def test_ukf_step_host_sees_a_changed_Q_and_R():
    """The zero-copy host step uploads the shared Q / R only when their values change; a change must take effect."""
    import torch
    B = 4096
    sc = synth.ukfom_scenario(B, seed=93)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    hu, hz = pin(sc["u"]), pin(sc["z"])
    hQ, hR = pin(sc["Q"]), pin(sc["R"])
    a, b = engine.Ukf(B), engine.Ukf(B)
    for f in (a, b):
        f.set_state(sc["mu"], sc["P"])
    out = torch.empty((B, 10), dtype=torch.float64).pin_memory()
    for k in range(3):
        Q, R = sc["Q"] * (1.0 + k), sc["R"] * (1.0 + 0.5 * k)
        hQ.copy_(torch.from_numpy(Q))                       # same host buffers, new values
        hR.copy_(torch.from_numpy(R))
        a.step(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, sc["u"], sc["dt"], Q, sc["z"], R)
        b.step_host(engine.PM_UKFOM_IMU, engine.MM_GPS_POS, hu, sc["dt"], hQ, hz, hR, mu_out=out)
        np.testing.assert_array_equal(out.numpy(), a.mu())
    np.testing.assert_array_equal(a.P(), b.P())
